#!/bin/bash
# cycle accounting of the sequential rounds at SF10 / SF9 (thread 0's view), one channel
LD_LIBRARY_PATH=$PWD/lora-sdr_b200/variants/rxprof:$LD_LIBRARY_PATH python tools/block_bench.py --sf 10 --packets 16 --repeat 2 2>&1 | grep -E "RXPROF" | tail -4 | cut -c1-300
LD_LIBRARY_PATH=$PWD/lora-sdr_b200/variants/rxprof:$LD_LIBRARY_PATH python tools/block_bench.py --sf 9 --packets 16 --repeat 2 2>&1 | grep -E "RXPROF" | tail -4 | cut -c1-300
