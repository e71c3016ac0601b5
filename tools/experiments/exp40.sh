#!/bin/bash
python -m pytest tests/test_gpu_rx.py tests/test_loopback.py tests/test_gpu_blocks.py -x -q -m gpu 2>&1 | tail -1
python tools/block_bench.py --sf 7 9 10 12 --packets 16 --repeat 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],round(d['reference_block']['Msamples_per_s'],1),round(d['b200_block']['Msamples_per_s'],1),round(d['speedup'],2),d['identical'])"
python tools/rx_bench.py --channels 256 --sf 10 12 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],d['channels'],round(d['gpu_ms'],2),'ms',round(d['gpu_aggregate_Msamples_per_s']/1e3,1),'Gs/s',d['identical_to_reference_block'])"
