#!/bin/bash
mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for f in bench_k20 bench_default bench_reference; do cut -c1-160 gpurun_out/$f.json; done
