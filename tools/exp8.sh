#!/bin/bash
L=lora-sdr_b200/liblorab200.so
tools/kbench --lib $L --tag idle0 --sf 7 --windows 65536 --steps 20 --reps 9 --print-reps
tools/kbench --lib $L --tag idle1 --sf 7 --windows 65536 --steps 20 --reps 9 --print-reps --idle-ms 1
tools/kbench --lib $L --tag idle20 --sf 7 --windows 65536 --steps 20 --reps 9 --print-reps --idle-ms 20
tools/kbench --lib $L --tag idle200 --sf 7 --windows 65536 --steps 20 --reps 5 --print-reps --idle-ms 200
