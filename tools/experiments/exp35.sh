#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_acquire.py -x -q -m gpu 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err; tail -2 gpurun_out/bench_k20.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_k20.json','gpurun_out/bench_default.json'):
    d=json.load(open(f)); e=d['e2e']
    print(f, round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', round(e['value'],2), e['repetitions'], 'pageable', round(e['pageable_input_value'],2))
    print('   ', ' '.join('%s %.4f'%(s_['label'],s_['roofline']['frac']) for s_ in d['sweep']), 'cfg4 %.4f'%d['cfg4']['roofline']['frac'])
PY
