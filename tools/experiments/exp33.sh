#!/bin/bash
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-sweep 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('value',round(d['value'],1),round(d['roofline']['frac'],4),'e2e',round(e['value'],2),e['repetitions'],'pageable',round(e['pageable_input_value'],2),'copy alone',round(e['h2d_copy_alone_gbs_per_gpu'],1),'cpu',round(d['cpu_baseline']['value'],1))"
done
python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | cut -c1-200
