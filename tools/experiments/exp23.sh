#!/bin/bash
# multi-rank bench line (usage: exp23.sh N)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',d['ms_per_step'],'frac',round(d['roofline']['frac'],4),'n_gpus',d['n_gpus'])
e=d['e2e']; print('e2e',round(e['value'],2),'pageable',round(e['pageable_input_value'],2),'copy alone',round(e['h2d_copy_alone_gbs_per_gpu'],1),'in e2e',round(e['h2d_gbs_in_e2e_per_gpu'],1),e['check'])
print('placement',d['host_placement'])
for s in d['sweep']: print(s['label'],round(s['value'],1),round(s['roofline']['frac'],4))
print('cfg4',round(d['cfg4']['value'],1),d['cfg4']['channels_per_gpu'],round(d['cfg4']['roofline']['frac'],4),d['cfg4']['indices_bit_exact'])
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['one_thread']['value'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_reference_n$N.json 2>/dev/null; cut -c1-260 gpurun_out/bench_reference_n$N.json
