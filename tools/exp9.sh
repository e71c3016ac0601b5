#!/bin/bash
mkdir -p gpurun_out
L=lora-sdr_b200/liblorab200.so
for sf in 7 10 12; do
w=$([ $sf = 7 ] && echo 65536 || echo 16384)
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 24 --csv --log-file gpurun_out/walk_launches_sf$sf.csv tools/kbench --lib $L --sf $sf --windows $w --steps 4 --reps 1 --fine-step 1.28 > /dev/null 2>&1
python3 - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/walk_launches_sf$sf.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:14]: print('sf$sf', r[4][:60], r[-1])
PY
done
tools/kbench --lib $L --sf 7 --windows 65536 --steps 100 --fine-step 1.28
tools/kbench --lib $L --sf 7 --windows 65536 --steps 100 --fine-step 0
