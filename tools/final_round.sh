#!/bin/bash
# Run ON THE GPU BOX (through gpurun): everything the round's profiles/ entries are made from.
# Outputs into gpurun_out/; tools/summarize_ncu.py + a manual copy turn them into profiles/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv,noheader > gpurun_out/gpu_info.txt
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
L=lora-sdr_b200/liblorab200.so
O=gpurun_out/kbench_table.jsonl; : > $O
tools/kbench --lib $L --tag dense --sf 7 --windows 16384,65536,524288 --steps 200 >> $O
tools/kbench --lib $L --tag dense --sf 8,9 --windows 16384,65536 --steps 100 >> $O
tools/kbench --lib $L --tag dense --sf 10,11,12 --windows 16384,65536 --steps 40 >> $O
tools/kbench --lib $L --tag not_input_stable --flags 0 --sf 7,10,12 --windows 65536 --steps 40 >> $O
tools/kbench --lib $L --tag k20 --sf 7 --windows 65536 --steps 20 --reps 9 >> $O
for fs in 0 1.28 13.37; do
  tools/kbench --lib $L --tag fine$fs --sf 7 --windows 65536 --steps 100 --fine-step $fs >> $O
  tools/kbench --lib $L --tag fine$fs --sf 10,12 --windows 16384 --steps 40 --fine-step $fs >> $O
done
python tools/strawman_cufft.py --sf 7 --windows 65536 --steps 100 > gpurun_out/strawman_sf7.json 2>/dev/null
python tools/block_bench.py --sf 7 10 12 --packets 16 --repeat 4 > gpurun_out/block_bench_thresh-30.jsonl 2>/dev/null
python tools/block_bench.py --sf 7 10 12 --packets 16 --repeat 4 --thresh -5 > gpurun_out/block_bench_thresh-5.jsonl 2>/dev/null
python tools/rx_bench.py --channels 256 > gpurun_out/rx_bench_256.jsonl 2>/dev/null
python tools/rx_bench.py --channels 1024 --sf 7 10 > gpurun_out/rx_bench_1024.jsonl 2>/dev/null
python tools/acquire_bench.py > gpurun_out/acquire_bench.jsonl 2>/dev/null
python tools/mod_bench.py > gpurun_out/mod_bench.jsonl 2> gpurun_out/mod_bench.err
python tools/call_latency.py > gpurun_out/call_latency.txt 2>&1
compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/racecheck.log 2>&1; tail -2 gpurun_out/racecheck.log
compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/memcheck.log 2>&1; tail -2 gpurun_out/memcheck.log
# (the five --set full reports together exceed gpurun's 64 MiB return limit: SKIP_NCU=1 runs the pass without
#  them, `bash tools/profile.sh` is then a gpurun call of its own)
[ -n "${SKIP_NCU:-}" ] || bash tools/profile.sh
for f in bench_k20 bench_default bench_reference; do cut -c1-200 gpurun_out/$f.json; done
