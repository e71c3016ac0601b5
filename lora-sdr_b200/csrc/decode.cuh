// decode.cuh -- batched symbols -> bytes decoder (row f4): one packet per thread running
// decode_packet (codec.h, the restatement of LoRaDecoder.cpp:196-397).  Integer work on a few
// hundred bytes per packet: only worth a launch when many packets (channels) are decoded at
// once, which is what the multi-channel receiver produces.
#pragma once
#include "codec.h"
#include "common.cuh"

namespace lorab200 {

struct DecodeParams {
    const uint16_t *symbols;   // [n_packets][sym_stride]
    unsigned long long n_packets;
    unsigned int sym_stride;
    const int32_t *n_syms;     // per-packet symbol counts, or nullptr: sym_stride symbols each
    DecoderCfg cfg;
    uint8_t *out;              // [n_packets][out_stride]
    unsigned int out_stride;
    int32_t *lengths;          // >= 0 bytes written; kDecodeDropped / kDecodeIgnored / kDecodeNoRoom
    // per-packet scratch
    uint16_t *sym_scratch; unsigned int max_sym;
    uint8_t *cw_scratch; unsigned int max_cw;
    uint8_t *byte_scratch; unsigned int max_bytes;
};

__global__ void decode_kernel(const DecodeParams p)
{
    const unsigned long long pk = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pk >= p.n_packets) return;
    unsigned n_in = p.sym_stride;
    if (p.n_syms) {
        const int32_t v = p.n_syms[pk];
        n_in = v < 0 ? 0u : (unsigned(v) > p.sym_stride ? p.sym_stride : unsigned(v));
    }
    uint16_t *sym = p.sym_scratch + pk * p.max_sym;
    const uint16_t *src = p.symbols + pk * p.sym_stride;
    for (unsigned i = 0; i < n_in; i++) sym[i] = src[i];
    p.lengths[pk] = decode_packet(p.cfg, sym, n_in, p.cw_scratch + pk * p.max_cw, p.byte_scratch + pk * p.max_bytes,
                                  p.out + pk * p.out_stride, p.out_stride);
}

} // namespace lorab200
