#!/bin/bash
# extra tiles spread evenly over the grid (new) against extras on the first groups (base)
mkdir -p gpurun_out
for lib in lora-sdr_b200/variants/liblorab200_base.so lora-sdr_b200/liblorab200.so; do
  tools/kbench --lib $lib --sf 7,8,9,10,11,12 --windows 16384 --steps 100 --tag $(basename $(dirname $lib))
  tools/kbench --lib $lib --sf 7,10 --windows 65536 --steps 100 --tag $(basename $(dirname $lib))
done > gpurun_out/exp36_kbench.jsonl 2> gpurun_out/exp36_kbench.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/exp36_kbench.jsonl')]
by={}
for r in rows: by.setdefault((r['sf'],r['windows']),{})[r['tag']]=r
for k in sorted(by):
    b=by[k]['variants']
    print(k, ' | '.join('%s %.3f us %.4f %s' % (t, v['us_per_step'], v['frac'], 'ok' if v['hash']==b['hash'] else 'HASH!') for t,v in by[k].items()))
PY
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -1
