#!/bin/bash
# uneven tile split between co-resident CTAs (LORAB200_SKEW) against the even split
mkdir -p gpurun_out
for lib in base skew10 skew20 skew30; do
  tools/kbench --lib lora-sdr_b200/variants/liblorab200_$lib.so --sf 7,10,12 --windows 16384,65536 --steps 100 --tag $lib
  tools/kbench --lib lora-sdr_b200/variants/liblorab200_$lib.so --sf 8,9,11 --windows 16384 --steps 100 --tag $lib
done > gpurun_out/exp19_kbench.jsonl 2> gpurun_out/exp19_kbench.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/exp19_kbench.jsonl')]
by={}
for r in rows: by.setdefault((r['sf'],r['windows']),{})[r['tag']]=r
for k in sorted(by):
    b=by[k]['base']
    print(k, ' | '.join('%s %.3f us %.4f %s' % (t, v['us_per_step'], v['frac'], 'ok' if v['hash']==b['hash'] else 'HASH!') for t,v in by[k].items()))
PY
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
for t in 8 16; do
LORAB200_STAGE_THREADS=$t python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage threads $t: value',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value'],2),'pageable',round(d['e2e']['pageable_input_value'],2),d['e2e']['check'])
for s_ in d['sweep']: print('  ',s_['label'],round(s_['value'],1),round(s_['roofline']['frac'],4))
"
done
LORAB200_NO_STREAM_COPY=1 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('memcpy fill: pageable',round(d['e2e']['pageable_input_value'],2))"
for lib in lora-sdr_b200/variants/liblorab200_base.so lora-sdr_b200/liblorab200.so; do
  tools/kbench --lib $lib --sf 7,10,12 --windows 16384 --steps 100 --fine-step 1.28 --tag walk1.28 | cut -c1-200
  tools/kbench --lib $lib --sf 7,10 --windows 16384 --steps 100 --fine-step 13.37 --tag walk13.37 | cut -c1-200
done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fine or walk" 2>&1 | tail -2
