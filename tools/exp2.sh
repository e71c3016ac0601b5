#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/exp2.jsonl; : > $O
V=lora-sdr_b200/variants
for a in base cta1 cta3 k3 k4; do
  tools/kbench --lib $V/liblorab200_$a.so --tag $a --sf 7 --windows 65536,524288 --steps 200 >> $O
  tools/kbench --lib $V/liblorab200_$a.so --tag $a --sf 8 --windows 16384,262144 --steps 100 >> $O
  tools/kbench --lib $V/liblorab200_$a.so --tag $a --sf 10,12 --windows 16384 --steps 50 >> $O
done
cat $O
timeout 300 ncu --set full --import-source on --clock-control none -k regex:demod_fused -c 1 -f -o gpurun_out/prof_r2_sf7_1m tools/kbench --lib $V/liblorab200_base.so --sf 7 --windows 1048576 --steps 2 --reps 1 > gpurun_out/ncu_r2_sf7_1m.log 2>&1
tail -3 gpurun_out/ncu_r2_sf7_1m.log
