#!/bin/bash
set -u
for K in 20 20 2000 20; do
python bench.py --no-sweep --no-cpu-baseline --e2e-steps 3 --steps $K --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('K=$K', round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), d['clocks'], 'pageable', round(d['e2e']['pageable_input_value'],1))"
done
