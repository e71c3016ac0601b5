#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rx.py tests/test_loopback.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/rx_bench.py --channels 256 2>&1 | tail -8 | cut -c60-330
timeout 300 python tools/rx_bench.py --channels 1024 --sf 7 10 2>&1 | tail -8 | cut -c60-330
timeout 300 python tools/block_bench.py --sf 7 10 12 --packets 16 --repeat 4 2>&1 | tail -4 | cut -c1-700
timeout 300 python tools/block_bench.py --sf 7 10 12 --packets 16 --repeat 4 --thresh -5 2>&1 | tail -4 | cut -c1-700
