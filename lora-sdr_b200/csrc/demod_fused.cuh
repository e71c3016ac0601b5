// placeholder until the fused persistent kernel lands (next commit)
#pragma once
#include "common.cuh"
namespace lorab200 {
inline bool fused_applicable(int, const DemodParams &) { return false; }
inline cudaError_t launch_fused(int, const DemodParams &, cudaStream_t, int) { return cudaErrorNotSupported; }
}
