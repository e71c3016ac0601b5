#!/bin/bash
# first half of the grid heavier by SKEW percent (the direction the plain "extras in front" split already leans)
mkdir -p gpurun_out
for lib in base hv4 hv8 hv16; do
  tools/kbench --lib lora-sdr_b200/variants/liblorab200_$lib.so --sf 7 --windows 65536 --steps 100 --tag $lib
  tools/kbench --lib lora-sdr_b200/variants/liblorab200_$lib.so --sf 9,10,12 --windows 16384 --steps 100 --tag $lib
done > gpurun_out/exp37_kbench.jsonl 2> gpurun_out/exp37_kbench.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/exp37_kbench.jsonl')]
by={}
for r in rows: by.setdefault((r['sf'],r['windows']),{})[r['tag']]=r
for k in sorted(by):
    b=by[k]['base']
    print(k, ' | '.join('%s %.3f us %.4f %s' % (t, v['us_per_step'], v['frac'], 'ok' if v['hash']==b['hash'] else 'HASH!') for t,v in by[k].items()))
PY
