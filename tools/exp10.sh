#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "fine or walk" 2>&1 | tail -2
L=lora-sdr_b200/liblorab200.so
for fs in 1.28 13.37; do
tools/kbench --lib $L --tag walk$fs --sf 7 --windows 65536 --steps 100 --fine-step $fs
tools/kbench --lib $L --tag walk$fs --sf 8,9 --windows 16384 --steps 100 --fine-step $fs
tools/kbench --lib $L --tag walk$fs --sf 10,11,12 --windows 16384 --steps 40 --fine-step $fs
done
