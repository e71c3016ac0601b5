#!/bin/bash
# Run ON THE GPU BOX (through gpurun): everything the round's profiles/ entries are made from.
# Outputs into gpurun_out/; tools/summarize_ncu.py + a manual copy turn them into profiles/.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for sf in 8 9 10 11 12; do
  python bench.py --sf $sf --windows 16384 --steps 400 --warmup 10 > gpurun_out/bench_sf$sf.json 2> gpurun_out/bench_sf$sf.err
done
python bench.py --no-input-stable --steps 1000 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/bench_sf7_not_stable.json 2>/dev/null
python bench.py --sf 10 --windows 1024 --channels 32 --steps 400 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/bench_cfg4_32ch.json 2>/dev/null
for sf in 7 10 12; do
  w=$([ $sf = 7 ] && echo 65536 || echo 16384)
  python bench.py --sf $sf --windows $w --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 3 --fine-step 13.37 > gpurun_out/bench_walk_sf$sf.json 2>/dev/null
  python bench.py --sf $sf --windows $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 --fine-step 13.37 --simple > gpurun_out/bench_walk_simple_sf$sf.json 2>/dev/null
  python tools/strawman_cufft.py --sf $sf --windows $w --steps 100 > gpurun_out/strawman_sf$sf.json 2>/dev/null
done
python tools/block_bench.py --sf 7 10 12 --packets 16 --repeat 4 > gpurun_out/block_bench_thresh-30.jsonl 2>/dev/null
python tools/block_bench.py --sf 7 10 12 --packets 16 --repeat 4 --thresh -5 > gpurun_out/block_bench_thresh-5.jsonl 2>/dev/null
python tools/mod_bench.py > gpurun_out/mod_bench.jsonl 2> gpurun_out/mod_bench.err
python tools/call_latency.py > gpurun_out/call_latency.txt 2>&1
python tools/rx_bench.py --channels 256 > gpurun_out/rx_bench_256.jsonl 2>/dev/null
python tools/rx_bench.py --channels 1024 > gpurun_out/rx_bench_1024.jsonl 2>/dev/null
bash tools/profile.sh
for f in bench_default bench_reference; do cut -c1-300 gpurun_out/$f.json; done
