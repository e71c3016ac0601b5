#!/bin/bash
python -m pytest tests/test_gpu_acquire.py -x -q -m gpu 2>&1 | tail -30
