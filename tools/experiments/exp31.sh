#!/bin/bash
mkdir -p gpurun_out
python tools/sanitize_small.py 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:rx_batch -s 2 -c 1 -f -o gpurun_out/prof_rx_sf7 \
    python tools/rx_bench.py --channels 256 --sf 7 > gpurun_out/ncu_rx.log 2>&1; tail -2 gpurun_out/ncu_rx.log | cut -c1-200
