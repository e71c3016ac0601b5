#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/exp3.jsonl; : > $O
V=lora-sdr_b200/variants
for a in "$@"; do
  tools/kbench --lib $V/liblorab200_$a.so --tag $a --sf 7 --windows 65536,524288 --steps 200 >> $O
  tools/kbench --lib $V/liblorab200_$a.so --tag $a --sf 8 --windows 16384,262144 --steps 100 >> $O
  tools/kbench --lib $V/liblorab200_$a.so --tag $a --sf 9 --windows 16384 --steps 100 >> $O
  tools/kbench --lib $V/liblorab200_$a.so --tag $a --sf 10,11,12 --windows 16384 --steps 50 >> $O
done
cat $O
