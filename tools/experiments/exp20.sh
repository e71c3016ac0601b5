#!/bin/bash
# full GPU test suite + smoke + short bench with the current build
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python tools/rx_bench.py --channels 256 --sf 7 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],d['channels'],round(d['gpu_aggregate_Msamples_per_s']/1e3,1),'Gs/s',d['identical_to_reference_block'])"
python tools/block_bench.py --sf 7 8 --packets 16 --repeat 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],round(d['reference_block']['Msamples_per_s'],1),round(d['b200_block']['Msamples_per_s'],1),round(d['speedup'],2),d['identical'])"
