#!/bin/bash
set -u
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Thread|Core|Socket|NUMA" | head
LORAB200_TRACE=1 python - <<'PY' 2>&1 | tail -40
import numpy as np, time, sys
sys.path.insert(0,'.')
import lora_sdr_b200 as m
sf=7; nw=65536; n=128
d=m.Demod(sf)
rng=np.random.default_rng(1)
iq=(rng.standard_normal((nw,n),dtype=np.float32)+1j*rng.standard_normal((nw,n),dtype=np.float32)).astype(np.complex64)
out=np.empty(nw,m.RESULT_DTYPE)
for i in range(3):
    t0=time.perf_counter(); d.demod_host(iq,nw,out=out); dt=time.perf_counter()-t0
    print(f"pageable step {i}: {dt*1e3:.2f} ms -> {nw/dt/1e6:.1f} Msym/s", flush=True)
PY
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_k20b.json 2> gpurun_out/bench_k20b.err; tail -3 gpurun_out/bench_k20b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k20b.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'clocks',d['clocks'])
print('sustained',d['sustained'])
print('e2e',d['e2e']['value'],d['e2e']['pageable_input_value'])
PY
