#!/bin/bash
# tapered chunk tail in lorab200_demod_host: parity + e2e
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-sweep 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('e2e',round(e['value'],2),'pageable',round(e['pageable_input_value'],2),'copy alone',round(e['h2d_copy_alone_gbs_per_gpu'],1),'in e2e',round(e['h2d_gbs_in_e2e_per_gpu'],1),e['check'])"
done
python bench.py --steps 20 --warmup 5 --no-sweep --sf 10 --windows 16384 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('sf10 e2e',round(e['value'],2),'pageable',round(e['pageable_input_value'],2),'copy alone',round(e['h2d_copy_alone_gbs_per_gpu'],1),'in e2e',round(e['h2d_gbs_in_e2e_per_gpu'],1),e['check'])"
