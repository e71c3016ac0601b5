#!/bin/bash
# Run ON THE GPU BOX (through gpurun): ncu captures the judge wants, into gpurun_out/.
#   1. launch list of the default bench command (every launch with its device time)
#   2. one --set full capture of the fused kernel for SF 7 / 10 / 12
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sf7.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/launches_sf7.log 2>&1
for sf in 7 10 12; do
  w=$([ $sf = 7 ] && echo 65536 || echo 16384)
  ncu --set full --clock-control none --import-source on -k regex:demod_fused -s 8 -c 1 -f -o gpurun_out/prof_sf$sf \
      python bench.py --sf $sf --windows $w --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_sf$sf.log 2>&1
  tail -1 gpurun_out/ncu_sf$sf.log | cut -c1-80
done
