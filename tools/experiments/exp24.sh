#!/bin/bash
# alternating uneven tile split (LORAB200_SKEW with launch parity) against the even split
mkdir -p gpurun_out
for lib in base alt10 alt20 alt40; do
  tools/kbench --lib lora-sdr_b200/variants/liblorab200_$lib.so --sf 7 --windows 65536 --steps 100 --tag $lib
  tools/kbench --lib lora-sdr_b200/variants/liblorab200_$lib.so --sf 9,10,12 --windows 16384 --steps 100 --tag $lib
done > gpurun_out/exp24_kbench.jsonl 2> gpurun_out/exp24_kbench.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/exp24_kbench.jsonl')]
by={}
for r in rows: by.setdefault((r['sf'],r['windows']),{})[r['tag']]=r
for k in sorted(by):
    b=by[k]['base']
    print(k, ' | '.join('%s %.3f us %.4f %s' % (t, v['us_per_step'], v['frac'], 'ok' if v['hash']==b['hash'] else 'HASH!') for t,v in by[k].items()))
PY
