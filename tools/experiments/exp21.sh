#!/bin/bash
# is the many-channel receiver time stable across process histories? (carve-out pinned)
for i in 1 2 3; do
python tools/rx_bench.py --channels 256 --sf 7 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('run $i',d['sf'],d['thresh_db'],d['channels'],round(d['gpu_ms'],2),'ms',round(d['gpu_aggregate_Msamples_per_s']/1e3,1),'Gs/s',d['identical_to_reference_block'])"
python tools/block_bench.py --sf 7 --packets 16 --repeat 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],round(d['reference_block']['Msamples_per_s'],1),round(d['b200_block']['Msamples_per_s'],1),round(d['speedup'],2),d['identical'])"
done
python tools/rx_bench.py --channels 1024 --sf 7 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],d['channels'],round(d['gpu_ms'],2),'ms',round(d['gpu_aggregate_Msamples_per_s']/1e3,1),'Gs/s',d['identical_to_reference_block'])"
python bench.py --steps 20 --warmup 5 --no-sweep 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['e2e']))"
