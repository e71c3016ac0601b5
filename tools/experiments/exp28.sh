#!/bin/bash
python tools/acquire_bench.py 2>&1 | tail -8
