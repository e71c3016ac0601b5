#!/bin/bash
# the receiver at BASELINE config 4's per-GPU share: 32 channels x SF10
mkdir -p gpurun_out
python tools/rx_bench.py --channels 32 --sf 10 > gpurun_out/rx_bench_32.jsonl 2>/dev/null
cat gpurun_out/rx_bench_32.jsonl | cut -c1-400
