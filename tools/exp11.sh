#!/bin/bash
mkdir -p gpurun_out
L=lora-sdr_b200/liblorab200.so
timeout 300 ncu --set full --import-source on --clock-control none -k regex:demod_fused -c 1 -f -o gpurun_out/prof_r2_orbit_sf7 tools/kbench --lib $L --sf 7 --windows 65536 --steps 2 --reps 1 --fine-step 1.28 > gpurun_out/ncu_orbit.log 2>&1
tail -2 gpurun_out/ncu_orbit.log
