#!/bin/bash
# GPU box: the parity tests and the driver-style bench lines (both arms)
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_k20.json 2> gpurun_out/bench_ref_k20.err; tail -c 600 gpurun_out/bench_ref_k20.json; tail -3 gpurun_out/bench_ref_k20.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err; tail -5 gpurun_out/bench_k20.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -5 gpurun_out/bench_default.err
