#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...] -> lora-sdr_b200/variants/liblorab200_NAME.so (kernel-tuning experiments; kbench --lib)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p lora-sdr_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared "$@" \
  -o lora-sdr_b200/variants/liblorab200_$name.so lora-sdr_b200/csrc/lorab200.cu lora-sdr_b200/csrc/finalize.cpp lora-sdr_b200/csrc/fine_walk.cpp lora-sdr_b200/csrc/tables.cpp
echo built $name
