#!/bin/bash
# GPU box: W sweep + timing ablations of the fused kernel (kbench)
set -u
mkdir -p gpurun_out
O=gpurun_out/exp1.jsonl; : > $O
V=lora-sdr_b200/variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> gpurun_out/exp1.info
tools/kbench --lib $V/liblorab200_base.so --tag base --sf 7 --windows 8192,16384,32768,65536,131072,262144,1048576 --steps 200 >> $O
tools/kbench --lib $V/liblorab200_base.so --tag base --sf 8,9 --windows 16384,65536 --steps 200 >> $O
tools/kbench --lib $V/liblorab200_base.so --tag base --sf 10 --windows 4096,16384,65536 --steps 100 >> $O
tools/kbench --lib $V/liblorab200_base.so --tag base --sf 11,12 --windows 4096,16384,32768 --steps 50 >> $O
for a in abl1 abl2 abl4 abl8 abl15; do
  tools/kbench --lib $V/liblorab200_$a.so --tag $a --sf 7 --windows 65536,524288 --steps 200 >> $O
  tools/kbench --lib $V/liblorab200_$a.so --tag $a --sf 10,12 --windows 16384 --steps 50 >> $O
done
tools/kbench --lib $V/liblorab200_base.so --tag base_nostable --flags 0 --sf 7 --windows 65536 --steps 200 >> $O
cat $O
