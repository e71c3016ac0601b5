#!/bin/bash
mkdir -p gpurun_out
python tools/block_bench.py --sf 7 10 12 --packets 16 --repeat 4 > gpurun_out/block_bench_thresh-30.jsonl 2>/dev/null
python tools/block_bench.py --sf 7 10 12 --packets 16 --repeat 4 --thresh -5 > gpurun_out/block_bench_thresh-5.jsonl 2>/dev/null
cat gpurun_out/block_bench_thresh-30.jsonl gpurun_out/block_bench_thresh-5.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],round(d['reference_block']['Msamples_per_s'],1),round(d['b200_block']['Msamples_per_s'],1),round(d['speedup'],2),d['identical'])"
