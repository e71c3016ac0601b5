#!/bin/bash
# receiver with the 32-bit index walk + planned window starts: parity tests, cycle accounting, block bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rx.py tests/test_loopback.py tests/test_gpu_blocks.py -x -q -m gpu 2>&1 | tail -3
LD_LIBRARY_PATH=$PWD/lora-sdr_b200/variants/rxprof:$LD_LIBRARY_PATH python tools/block_bench.py --sf 7 --packets 16 --repeat 4 2>&1 | grep -E "RXPROF" | tail -3 | cut -c1-300
python tools/block_bench.py --sf 7 8 9 10 12 --packets 16 --repeat 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],round(d['reference_block']['Msamples_per_s'],1),round(d['b200_block']['Msamples_per_s'],1),round(d['speedup'],2),d['identical'])"
python tools/rx_bench.py --channels 256 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['sf'],d['thresh_db'],d['channels'],round(d['gpu_aggregate_Msamples_per_s']/1e3,1),'Gs/s per-ch',round(d['gpu_per_channel_Msamples_per_s'],1),'ref1thr',round(d['reference_block_1thread_Msamples_per_s'],1),d['identical_to_reference_block'])"
