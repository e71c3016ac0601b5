// mod_synth.cuh -- on-GPU chirp synthesis / modulator.
//
// Replaces, for batches of packets, LoRaMod.cpp:120-229 (frame state machine: 10 preamble
// up-chirps, two sync-word chirps, two down-chirps and a quarter down-chirp, the data
// symbols, zero padding) with genChirp<float> (ChirpGenerator.hpp:21-47) as the signal law.
//
// Two modes, because the reference's signal law is a pair of sequential single-precision
// recurrences (f += fStep; phaseAccum += f) whose rounding errors pile up (at SF12 the
// reference's own waveform drifts from the ideal chirp by more than a radian over a packet):
//
//   LORAB200_MOD_REFERENCE  the recurrences are evaluated literally, one thread per packet
//                           (they are sequential across the whole packet: the phase carries
//                           from chirp to chirp), so the PHASE of every sample is bit-identical
//                           to the reference's; a second, fully parallel pass turns phases
//                           into samples.  polar() uses cos/sin evaluated in double and rounded
//                           to float where the reference calls glibc's cosf/sinf: both are
//                           within an ulp of the exact value, so samples agree to <= 2 ulp
//                           (tests/test_gpu_mod.py states and checks the bound).
//   LORAB200_MOD_IDEAL      closed form.  With ovs = 1 every phase of an ideal LoRa frame is an
//                           integer multiple of pi/N: sample i of a chirp that starts at
//                           phase index M has index M +- ((i+1)(2s - N) + (i+1)(i+2)) mod 2N.
//                           One table of 2N phasors, no recurrence: every sample independent,
//                           the kernel runs at HBM write speed.  Behavioural parity: the
//                           reference receiver decodes these frames (tests), sample values
//                           differ from the reference's by the reference's own drift.
#pragma once
#include "common.cuh"

namespace lorab200 {

struct ModParams {
    const uint16_t *symbols; // [n_packets][n_syms]
    unsigned long long n_packets;
    unsigned int n_syms;     // data symbols per packet
    unsigned int padding;    // zero symbols after the data (LoRaMod::setPadding)
    unsigned int sync;       // sync word (LoRaMod::setSync)
    unsigned int ovs;        // oversampling (LoRaMod::setOvs); IDEAL mode: 1
    unsigned int n;          // N = 2^SF
    float ampl;
    unsigned long long samples_per_packet;
    float2 *out;             // [n_packets][samples_per_packet]
    float *phase;            // REFERENCE mode scratch, same shape
    const float2 *phasor;    // IDEAL mode: exp(j*pi*m/N), m < 2N
};

constexpr int kModPreamble = 10; // LoRaMod.cpp:137  _counter = 10

// chirps in front of the data: preamble, 2 sync words, 2 down-chirps (+ the quarter chirp)
constexpr int kModHeadChirps = kModPreamble + 4;

__host__ __device__ inline unsigned long long mod_samples_per_packet(unsigned n, unsigned ovs, unsigned n_syms, unsigned padding)
{
    const unsigned long long nn = (unsigned long long)n * ovs;
    return (kModHeadChirps + (unsigned long long)n_syms + padding) * nn + nn / 4;
}

// ---- REFERENCE mode, pass 1: the phase recurrence, literally -------------------------------
// ChirpGenerator.hpp:22-47 with Type = float.  Returns the running phase accumulator.
__device__ __forceinline__ float gen_chirp_phase(float *ph, int N, int ovs, int NN, float f0, bool down, float phaseAccum)
{
    const float fMin = float(-M_PI / ovs);
    const float fMax = float(M_PI / ovs);
    const float fStep = float((2 * M_PI) / (double(N) * ovs * ovs));
    const float span = __fsub_rn(fMax, fMin);
    float f = __fadd_rn(fMin, f0);
    for (int i = 0; i < NN; i++) {
        f = __fadd_rn(f, fStep);
        if (f > fMax) f = __fsub_rn(f, span);
        phaseAccum = down ? __fsub_rn(phaseAccum, f) : __fadd_rn(phaseAccum, f);
        ph[i] = phaseAccum;
    }
    // phaseAccum -= floor(phaseAccum / (2 * M_PI)) * 2 * M_PI;   (double arithmetic, one rounding)
    const double w = __dmul_rn(__dmul_rn(floor(__ddiv_rn(double(phaseAccum), 2 * M_PI)), 2.0), M_PI);
    return __double2float_rn(__dsub_rn(double(phaseAccum), w));
}

__global__ void mod_phase_kernel(const ModParams p)
{
    const unsigned long long pk = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pk >= p.n_packets) return;
    const int N = int(p.n), ovs = int(p.ovs), NN = N * ovs;
    float *ph = p.phase + pk * p.samples_per_packet;
    const uint16_t *sym = p.symbols + pk * p.n_syms;
    float acc = 0.0f;                                        // LoRaMod.cpp:138  _phaseAccum = 0
    for (int c = 0; c < kModPreamble; ++c) { acc = gen_chirp_phase(ph, N, ovs, NN, 0.0f, false, acc); ph += NN; }
    const int sw0 = int(p.sync >> 4) * 8, sw1 = int(p.sync & 0xf) * 8;     // LoRaMod.cpp:157,168
    acc = gen_chirp_phase(ph, N, ovs, NN, float((2 * M_PI * sw0) / double(NN)), false, acc); ph += NN;
    acc = gen_chirp_phase(ph, N, ovs, NN, float((2 * M_PI * sw1) / double(NN)), false, acc); ph += NN;
    acc = gen_chirp_phase(ph, N, ovs, NN, 0.0f, true, acc); ph += NN;
    acc = gen_chirp_phase(ph, N, ovs, NN, 0.0f, true, acc); ph += NN;
    acc = gen_chirp_phase(ph, N, ovs, NN / 4, 0.0f, true, acc); ph += NN / 4;  // LoRaMod.cpp:197
    for (unsigned k = 0; k < p.n_syms; ++k) {
        const float freq = float((2 * M_PI * int(sym[k])) / double(NN));       // LoRaMod.cpp:207-208
        acc = gen_chirp_phase(ph, N, ovs, NN, freq, false, acc);
        ph += NN;
    }
}

// ---- REFERENCE mode, pass 2: samps[i] = std::polar(ampl, phase) ----------------------------
__global__ void mod_polar_kernel(const ModParams p)
{
    const unsigned long long total = p.n_packets * p.samples_per_packet;
    const unsigned long long live = p.samples_per_packet - (unsigned long long)p.padding * p.n * p.ovs;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const unsigned long long at = g % p.samples_per_packet;
        float2 v = make_float2(0.0f, 0.0f);                                   // LoRaMod.cpp:224  padding
        if (at < live) {
            double sn, cs;
            sincos(double(p.phase[g]), &sn, &cs);
            v.x = __fmul_rn(p.ampl, __double2float_rn(cs));
            v.y = __fmul_rn(p.ampl, __double2float_rn(sn));
        }
        p.out[g] = v;
    }
}

// ---- IDEAL mode: closed form ---------------------------------------------------------------
// phase index (units of pi/N) of sample i of a chirp with symbol s that starts at index M:
//   up:   M + (i+1)(2s - N) + (i+1)(i+2)      down: M - ((i+1)(-N) + (i+1)(i+2))      (mod 2N)
// a full chirp advances the index by N (mod 2N) whatever its symbol.
__device__ __forceinline__ unsigned mod_ideal_index(unsigned N, unsigned M, unsigned s, unsigned i, bool down)
{
    const unsigned m2 = 2 * N - 1;                      // 2N is a power of two
    const unsigned a = (i + 1) & m2;
    unsigned inc = (a * ((2 * s - N) & m2) + a * ((i + 2) & m2)) & m2;
    if (down) inc = (0u - ((a * ((0u - N) & m2) + a * ((i + 2) & m2)) & m2)) & m2;
    return (M + inc) & m2;
}

// One block = 512 consecutive samples of one packet (two per thread, one 16-byte store each);
// the only division is the block -> packet split, N is a power of two.
constexpr int kModIdealSpan = 512;

__device__ __forceinline__ float2 mod_ideal_sample(const ModParams &p, unsigned long long pk, unsigned at,
                                                   unsigned sf, unsigned head, unsigned Mq, unsigned sw0, unsigned sw1)
{
    const unsigned N = p.n;
    unsigned m;
    if (at < head) {
        const unsigned c = at >> sf, i = at & (N - 1);
        const bool down = c >= kModPreamble + 2;
        const unsigned s = c == kModPreamble ? sw0 : (c == kModPreamble + 1 ? sw1 : 0u);
        m = mod_ideal_index(N, (c & 1u) * N, s, i, down);
    } else if (at < head + N / 4) {
        m = mod_ideal_index(N, 0, 0, at - head, true);           // 14 full chirps: back at index 0
    } else {
        const unsigned d = at - head - N / 4;
        const unsigned k = d >> sf, i = d & (N - 1);
        if (k >= p.n_syms) return make_float2(0.0f, 0.0f);        // padding (LoRaMod.cpp:224)
        m = mod_ideal_index(N, Mq + (k & 1u) * N, p.symbols[pk * p.n_syms + k], i, false);
    }
    const float2 e = __ldg(p.phasor + m);
    return make_float2(__fmul_rn(p.ampl, e.x), __fmul_rn(p.ampl, e.y));
}

__global__ void __launch_bounds__(kModIdealSpan / 2) mod_ideal_kernel(const ModParams p, unsigned sf, unsigned blocks_per_packet)
{
    const unsigned N = p.n;
    const unsigned head = kModHeadChirps * N;            // samples before the quarter chirp
    const unsigned Mq = mod_ideal_index(N, 0, 0, N / 4 - 1, true); // index after the quarter chirp
    const unsigned sw0 = (p.sync >> 4) * 8, sw1 = (p.sync & 0xf) * 8;
    const unsigned long long n_blocks = p.n_packets * blocks_per_packet;
    for (unsigned long long b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const unsigned long long pk = b / blocks_per_packet;
        const unsigned at = unsigned(b - pk * blocks_per_packet) * kModIdealSpan + 2 * threadIdx.x;
        if (at >= p.samples_per_packet) continue;        // samples_per_packet is a multiple of 32
        const float2 a = mod_ideal_sample(p, pk, at, sf, head, Mq, sw0, sw1);
        const float2 c = mod_ideal_sample(p, pk, at + 1, sf, head, Mq, sw0, sw1);
        *reinterpret_cast<float4 *>(p.out + pk * p.samples_per_packet + at) = make_float4(a.x, a.y, c.x, c.y);
    }
}

} // namespace lorab200
