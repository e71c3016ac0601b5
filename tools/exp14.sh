#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'n_gpus',d['n_gpus'])
print('e2e',d['e2e']['value'],d['e2e']['pageable_input_value'],d['e2e']['check'])
print('placement',d['host_placement'])
for s in d['sweep']: print(s['label'],round(s['value'],1),round(s['roofline']['frac'],4))
print('cfg4',d['cfg4']['value'],d['cfg4']['channels_per_gpu'],d['cfg4']['roofline']['frac'],d['cfg4']['indices_bit_exact'])
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['one_thread']['value'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 | cut -c1-300
