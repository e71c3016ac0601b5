#!/bin/bash
# V1 (pass-0 scatter from one address register) against the previous build, all SFs; cycle accounting of the
# receiver's one-warp rounds at SF7 / -30 dB
mkdir -p gpurun_out
for lib in base v1; do
  tools/kbench --lib lora-sdr_b200/variants/liblorab200_$lib.so --sf 7,8,9,10,11,12 --windows 16384,65536 --steps 100 --tag $lib
done > gpurun_out/exp17_kbench.jsonl 2> gpurun_out/exp17_kbench.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/exp17_kbench.jsonl')]
by={}
for r in rows: by.setdefault((r['sf'],r['windows']),{})[r['tag']]=r
for k in sorted(by):
    b,v=by[k].get('base'),by[k].get('v1')
    if b and v: print(k, 'base %.3f us %.4f | v1 %.3f us %.4f | x%.3f | hash %s' % (b['us_per_step'],b['frac'],v['us_per_step'],v['frac'],b['us_per_step']/v['us_per_step'],'same' if b['hash']==v['hash'] else 'DIFFERENT'))
PY
LD_LIBRARY_PATH=$PWD/lora-sdr_b200/variants/rxprof:$LD_LIBRARY_PATH python tools/block_bench.py --sf 7 --packets 16 --repeat 4 2>&1 | grep -E "RXPROF|speedup" | tail -8 | cut -c1-400
