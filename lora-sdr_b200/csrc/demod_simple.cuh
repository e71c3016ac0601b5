// demod_simple.cuh -- the plain kernel: one CTA per symbol window, the window in
// shared memory, kissfft's in-place stages executed literally (one
// __syncthreads per stage).  It is the general-purpose path: any stride /
// alignment, the fine-tune recurrence with a non-zero step, the fft/dec debug
// dumps.  The fused persistent kernel (demod_fused.cuh) is the fast path and
// is cross-checked against this one on the GPU.
//
// Replaces, per window: LoRaDemod.cpp:157-166 (dechirp) + LoRaDetector.hpp:29-64
// (FFT + scan) with kissfft.hh:83-157 as the FFT.
#pragma once
#include "common.cuh"

namespace lorab200 {

// kissfft's load permutation for N = 2^SF (kissfft.hh:90-94,101-103): output
// slot j with mixed-radix digits (q_0 most significant) reads in[q_0 + r_0*(q_1 + ...)].
// Radices are 4,...,4 and a final 2 when SF is odd (kissfft.hh:34-51).
template <int SF>
__device__ __forceinline__ int kiss_input_index(int j)
{
    constexpr int L4 = SF / 2;      // number of radix-4 stages
    int i = 0, w = 1;
    // digit s (s = 0 is the most significant digit of j)
    int shift = SF;
#pragma unroll
    for (int s = 0; s < L4; ++s) {
        shift -= 2;
        i += ((j >> shift) & 3) * w;
        w <<= 2;
    }
    if (SF & 1) i += (j & 1) * w;
    return i;
}

// The inverse map: input sample i lands in slot j (i's base-4 digits, least significant first,
// are j's digits most significant first; the top bit of i is j's last bit when SF is odd).
template <int SF>
__device__ __forceinline__ int kiss_output_slot(int i)
{
    constexpr int L4 = SF / 2;
    int j = 0;
    int shift = SF;
#pragma unroll
    for (int s = 0; s < L4; ++s) {
        shift -= 2;
        j |= ((i >> (2 * s)) & 3) << shift;
    }
    if (SF & 1) j |= (i >> (2 * L4)) & 1;
    return j;
}

constexpr int kSimpleThreads = 128;

// One window, cooperatively by a CTA of kSimpleThreads threads: dechirp (+ optional fine-tune
// multiply: fidx[i] per sample when `walk`, else the constant index idx_const), kissfft's stages
// in shared memory, peak scan.  F: N float2 of shared memory; s_*: kSimpleThreads/32 entries each.
// The raw record is returned in thread 0 (other threads: unspecified).  Ends with a
// __syncthreads, so F and the scratch may be reused at once.
template <int SF>
__device__ __forceinline__ RawResult simple_window(const float2 *__restrict__ in, const float2 *__restrict__ chirp,
                                                   const float2 *__restrict__ fine, const int *fidx, int idx_const,
                                                   bool walk, const float2 *__restrict__ tw, float2 *dec_out,
                                                   float2 *fft_out, float2 *F, float *s_val, int *s_idx, double *s_sum)
{
    constexpr int N = 1 << SF;
    constexpr int T = kSimpleThreads;
    const int tid = threadIdx.x;

    // dechirp in input order (coalesced global reads), scatter into kissfft's digit-reversed slots
    for (int i = tid; i < N; i += T) {
        float2 d = cx_mul(in[i], chirp[i]);
        if (fine) d = cx_mul(d, fine[walk ? fidx[i] : idx_const]);
        if (dec_out) dec_out[i] = d;
        F[kiss_output_slot<SF>(i)] = d;
    }

    // stages, deepest first: radix-2 (odd SF) with m=1, then radix-4 with m=..,
    int m = 1;
    if (SF & 1) {
        __syncthreads();
        for (int b = tid; b < N / 2; b += T) {
            float2 a0 = F[2 * b], a1 = cx_mul(F[2 * b + 1], tw[0]);
            bfly2(a0, a1);
            F[2 * b] = a0; F[2 * b + 1] = a1;
        }
        m = 2;
    }
    for (; m < N; m <<= 2) {
        __syncthreads();
        const int fstride = N / (4 * m);
        for (int b = tid; b < N / 4; b += T) {
            const int k = b & (m - 1);
            const int base = (b - k) * 4 + k;
            float2 a0 = F[base];
            float2 a1 = cx_mul(F[base + m], tw[k * fstride]);
            float2 a2 = cx_mul(F[base + 2 * m], tw[2 * k * fstride]);
            float2 a3 = cx_mul(F[base + 3 * m], tw[3 * k * fstride]);
            bfly4(a0, a1, a2, a3);
            F[base] = a0; F[base + m] = a1; F[base + 2 * m] = a2; F[base + 3 * m] = a3;
        }
    }
    __syncthreads();

    // peak scan (LoRaDetector.hpp:33-48): strict '>' from 0, first max wins
    float bv = 0.0f; int bi = 0; double sum = 0.0;
    for (int j = tid; j < N; j += T) {
        const float2 v = F[j];
        if (fft_out) fft_out[j] = v;
        const float m2 = cx_mag2(v);
        sum += double(m2);
        if (m2 > bv) { bv = m2; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, bv, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        argmax_merge(bv, bi, ov, oi);
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = bv; s_idx[tid >> 5] = bi; s_sum[tid >> 5] = sum; }
    __syncthreads();
    RawResult r;
    r.index = 0; r.max_mag2 = 0.0f; r.total = 0.0; r.left = make_float2(0.f, 0.f); r.right = make_float2(0.f, 0.f);
    if (tid == 0) {
        for (int k = 1; k < T / 32; ++k) { argmax_merge(bv, bi, s_val[k], s_idx[k]); sum += s_sum[k]; }
        r.index = bi; r.max_mag2 = bv; r.total = sum;
        r.left = F[bi > 0 ? bi - 1 : N - 1];
        r.right = F[bi < N - 1 ? bi + 1 : 0];
    }
    __syncthreads();
    return r;
}

template <int SF, bool MULTI>
__global__ void __launch_bounds__(kSimpleThreads)
demod_simple_kernel(const DemodParams p, const SegList<MULTI> segs)
{
    constexpr int N = 1 << SF;
    constexpr int T = kSimpleThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *F = reinterpret_cast<float2 *>(smem_raw);          // N bins
    int *fidx = reinterpret_cast<int *>(F + N);                // N fine indices (only if needed)
    __shared__ float s_val[T / 32];
    __shared__ int s_idx[T / 32];
    __shared__ double s_sum[T / 32];

    const int tid = threadIdx.x;
    for (unsigned long long w = blockIdx.x; w < p.n_windows; w += gridDim.x) {
        const float2 *in;
        if constexpr (MULTI) {
            const unsigned long long c = w / p.seg_windows;
            in = segs.p[c] + (w - c * p.seg_windows) * p.stride;
        } else {
            in = p.iq + w * p.stride;
        }
        const bool fine = p.fine != nullptr;
        const bool walk = fine && p.fine_step != 0.0f;
        if (walk) {
            // sequential recurrence over the window (LoRaDemod.cpp:160-162)
            if (tid == 0) {
                int idx = p.fine_x16[w * (N / 16)];
                for (int i = 0; i < N; ++i) { fidx[i] = idx; idx = fine_next(idx, p.fine_step, p.fine_len); }
            }
            __syncthreads();
        }
        const RawResult r = simple_window<SF>(in, p.chirp, p.fine, fidx, p.fine_index0, walk, p.tw,
                                              p.dec ? p.dec + w * N : nullptr, p.fft ? p.fft + w * N : nullptr,
                                              F, s_val, s_idx, s_sum);
        if (tid == 0) {
            if (p.raw) p.raw[w] = r;
            if (p.out) finalize_device(r, p.power_scale, &p.out[w]);
        }
    }
}

} // namespace lorab200
