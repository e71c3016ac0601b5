#!/bin/bash
set -u
mkdir -p gpurun_out
L=lora-sdr_b200/liblorab200.so
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 50 > gpurun_out/exp5_clocks.csv &
SMI=$!
sleep 0.5
tools/kbench --lib $L --tag r --sf 7 --windows 65536 --steps 200 --reps 150 --print-reps
sleep 0.3
tools/kbench --lib $L --tag r --sf 7 --windows 65536 --steps 4000 --reps 6 --print-reps
kill $SMI
cat gpurun_out/exp5_clocks.csv | awk 'NR%2==1' | head -60
