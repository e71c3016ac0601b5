#!/bin/bash
# SF11 / SF12: chirp through L1 (LDG) instead of shared memory, which makes room for a second TMA stage
mkdir -p gpurun_out
for lib in base ldg48; do
  tools/kbench --lib lora-sdr_b200/variants/liblorab200_$lib.so --sf 11,12 --windows 16384,65536 --steps 60 --tag $lib
done > gpurun_out/exp30_kbench.jsonl 2> gpurun_out/exp30_kbench.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/exp30_kbench.jsonl')]
by={}
for r in rows: by.setdefault((r['sf'],r['windows']),{})[r['tag']]=r
for k in sorted(by):
    b=by[k]['base']
    print(k, ' | '.join('%s %.3f us %.4f %s' % (t, v['us_per_step'], v['frac'], 'ok' if v['hash']==b['hash'] else 'HASH!') for t,v in by[k].items()))
PY
