#!/bin/bash
# L2 prefetch of the look-ahead window in the CTA-wide sequential rounds (SF9-11): variant against the product
for v in product pfl2; do
  if [ $v = pfl2 ]; then export LD_LIBRARY_PATH=$PWD/lora-sdr_b200/variants/pfl2:$LD_LIBRARY_PATH; export LORAB200_LIB=$PWD/lora-sdr_b200/variants/pfl2/liblorab200.so; fi
  python tools/block_bench.py --sf 9 10 11 --packets 16 --repeat 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$v',d['sf'],d['thresh_db'],round(d['reference_block']['Msamples_per_s'],1),round(d['b200_block']['Msamples_per_s'],1),round(d['speedup'],2),d['identical'])"
  python tools/rx_bench.py --channels 256 --sf 9 10 11 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$v',d['sf'],d['thresh_db'],d['channels'],round(d['gpu_ms'],2),'ms',round(d['gpu_aggregate_Msamples_per_s']/1e3,1),'Gs/s',d['identical_to_reference_block'])"
done
