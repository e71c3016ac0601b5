#!/bin/bash
# one span copy per tile for overlapping windows: parity, dense timing unchanged?, search rate
python -m pytest tests/test_gpu_parity.py tests/test_gpu_acquire.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
for lib in lora-sdr_b200/variants/liblorab200_base.so lora-sdr_b200/liblorab200.so; do
  tools/kbench --lib $lib --sf 7,8 --windows 65536 --steps 100 --tag x | cut -c1-160
done
python tools/acquire_bench.py --sf 7 8 2>&1 | tail -4 | cut -c1-260
