#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/exp4.jsonl; : > $O
L=lora-sdr_b200/liblorab200.so
tools/kbench --lib $L --tag s20 --sf 7 --windows 65536 --steps 20 --reps 9 >> $O
tools/kbench --lib $L --tag s200 --sf 7 --windows 65536 --steps 200 >> $O
tools/kbench --lib $L --tag s2000 --sf 7 --windows 65536 --steps 2000 >> $O
tools/kbench --lib $L --tag s8000 --sf 7 --windows 65536 --steps 8000 --reps 3 >> $O
cat $O
LORAB200_NO_SAMPLER=1 python bench.py --no-sweep --no-cpu-baseline --e2e-steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nosampler K2000', d['ms_per_step'], d['roofline']['frac'])"
python bench.py --no-sweep --no-cpu-baseline --e2e-steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sampler K2000', d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
LORAB200_NO_SAMPLER=1 python bench.py --no-sweep --no-cpu-baseline --e2e-steps 3 --steps 200 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nosampler K200', d['ms_per_step'], d['roofline']['frac'])"
