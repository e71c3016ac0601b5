#!/bin/bash
python -m pytest tests/test_gpu_rx.py -x -q -m gpu 2>&1 | tail -5
