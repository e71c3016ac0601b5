#!/bin/bash
python -m pytest tests/test_gpu_rx.py tests/test_loopback.py tests/test_gpu_blocks.py tests/test_gpu_decode.py -x -q -m gpu 2>&1 | tail -2
python tools/rx_bench.py --channels 256 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],d['channels'],round(d['gpu_aggregate_Msamples_per_s']/1e3,1),'Gs/s per-ch',round(d['gpu_per_channel_Msamples_per_s'],1),'ref1thr',round(d['reference_block_1thread_Msamples_per_s'],1),d['identical_to_reference_block'])"
python tools/block_bench.py --sf 7 8 9 --packets 16 --repeat 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],round(d['reference_block']['Msamples_per_s'],1),round(d['b200_block']['Msamples_per_s'],1),round(d['speedup'],2),d['identical'])"
