#!/bin/bash
# timing ablations at SF10 / SF12 (results are wrong, timing only): what the peak scan, double sum and neighbour select cost
mkdir -p gpurun_out
for lib in base abl1 abl2 abl4 abl15; do
  tools/kbench --lib lora-sdr_b200/variants/liblorab200_$lib.so --sf 9,10,12 --windows 16384 --steps 100 --tag $lib
done > gpurun_out/exp26_kbench.jsonl 2> gpurun_out/exp26_kbench.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/exp26_kbench.jsonl')]
by={}
for r in rows: by.setdefault((r['sf'],r['windows']),{})[r['tag']]=r
for k in sorted(by):
    print(k, ' | '.join('%s %.3f us %.4f' % (t, v['us_per_step'], v['frac']) for t,v in by[k].items()))
PY
