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

// ---- encoder: payload bytes -> symbols, one packet per thread (encode_packet, codec.h) -------
struct EncodeParams {
    const uint8_t *payloads;   // [n_packets][payload_stride]
    unsigned long long n_packets;
    unsigned int payload_stride;
    const int32_t *lengths;    // per-packet payload lengths, or nullptr: payload_stride bytes each
    EncoderCfg cfg;
    uint16_t *symbols;         // [n_packets][sym_stride]
    unsigned int sym_stride;
    int32_t *n_symbols;        // symbols written per packet, or kDecodeNoRoom
    uint8_t *byte_scratch; unsigned int max_bytes;
    uint8_t *cw_scratch; unsigned int max_cw;
};

__global__ void encode_kernel(const EncodeParams p)
{
    const unsigned long long pk = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pk >= p.n_packets) return;
    unsigned len = p.payload_stride;
    if (p.lengths) {
        const int32_t v = p.lengths[pk];
        len = v < 0 ? 0u : (unsigned(v) > p.payload_stride ? p.payload_stride : unsigned(v));
    }
    if (codec_enc_symbols(p.cfg, len) > p.sym_stride) { p.n_symbols[pk] = kDecodeNoRoom; return; }
    p.n_symbols[pk] = encode_packet(p.cfg, p.payloads + pk * p.payload_stride, len, p.byte_scratch + pk * p.max_bytes,
                                    p.cw_scratch + pk * p.max_cw, p.symbols + pk * p.sym_stride);
}

} // namespace lorab200
