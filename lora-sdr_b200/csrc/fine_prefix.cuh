// fine_prefix.cuh -- the fine-tune index (LoRaDemod.cpp:160-162) at every 16th dechirped
// sample of a batch, computed on the device without walking the batch sequentially.
//
// The index is a one-word state x advanced once per sample by g(x) = wrap(trunc(fl32(
// float(x) - d))) (common.cuh: fine_next, the literal recurrence).  The demod kernels need
// the state at the start of every 16-sample chunk, X[q] = g^(16 q)(x0), q < W*N/16.
// g is a map on the finite set [0, N*128), so powers of g are lookup tables and the
// prefix is pointer jumping, radix R = 16:
//
//     T      = g^16                           one table entry = 16 literal steps
//     P_0    = T,   P_(i+1) = P_i^R           (table composition, N*128 entries)
//     X[m*b + r] = P_i^m [ X[r] ]             b = R^i, r < b, 1 <= m < R  (a chain of m lookups)
//
// ceil(log16(Q)) rounds of one small kernel each (5 for 2^19 chunks); everything is built
// from fine_next, so the result is exact by construction (tests compare with the sequential
// walk).  T only depends on (SF, d) and is cached in the context.
#pragma once
#include "common.cuh"

namespace lorab200 {

constexpr int kFineChunk = 16; // samples per prefix entry (= points per thread of the fused kernel)

__global__ void fine_base_kernel(int32_t *__restrict__ T, float step, int len)
{
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < len; x += gridDim.x * blockDim.x) {
        int y = x;
#pragma unroll
        for (int i = 0; i < kFineChunk; ++i) y = fine_next(y, step, len);
        T[x] = y;
    }
}

constexpr int kFineRadix = 16;

// One round with base b = R^i: thread r < b fills X[m*b + r], m = 1..R-1 (below q_end) by
// walking P from X[r]; when `compose` is set every thread also writes Pn = P^R.  Round 0
// (b == 1) stores X[0] = x0.
__global__ void fine_round_kernel(const int32_t *__restrict__ P, int32_t *__restrict__ Pn,
                                  int32_t *__restrict__ X, unsigned long long b,
                                  unsigned long long q_end, int len, int x0, int compose)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
         i < b || (compose && i < (unsigned long long)len); i += stride) {
        if (i < b) {
            int x = (b == 1) ? x0 : X[i];
            if (b == 1) X[0] = x0;
            for (int m = 1; m < kFineRadix; ++m) {
                const unsigned long long q = (unsigned long long)m * b + i;
                if (q >= q_end) break;
                x = P[x];
                X[q] = x;
            }
        }
        if (compose && i < (unsigned long long)len) {
            int y = int(i);
#pragma unroll
            for (int m = 0; m < kFineRadix; ++m) y = P[y];
            Pn[i] = y;
        }
    }
}

// ---- small batches: closed form per entry -----------------------------------------------
// A block's work() calls with a handful of windows; building g^16 over the whole table is
// then the dominant cost.  Instead every entry X[q] = g^(16 q)(x0) is evaluated independently
// with the region walk of fine_walk.h (host: FineWalk::advance, tested exhaustively against
// the sequential recurrence), restated for the device: a step is D+1 for 1 <= k < Kp, D for
// k >= Kp (k = x - D), everything else falls back to the literal recurrence.  D, Kp and the
// integer-step flag are computed once per call on the host (FineWalk's constructor).
struct FineWalkDev {
    int len;
    float d;
    int integer;
    long long D, Kp;
};

__device__ __forceinline__ int fine_advance_dev(const FineWalkDev &fw, int x, unsigned long long n)
{
    const long long L = fw.len;
    if (n == 0) return x;
    if (fw.integer) {
        long long m = ((long long)x - (long long)(n % (unsigned long long)L) * (fw.D % L)) % L;
        if (m < 0) m += L;
        return int(m);
    }
    const long long KP_NONE = 0x7fffffffffffffffll;
    while (n > 0) {
        const long long k = (long long)x - fw.D;
        unsigned long long j = 0;
        long long s = 0;
        if (k >= 1) {
            const bool r = (k >= fw.Kp);
            s = fw.D + 1 - (r ? 1 : 0);
            if (s == 0) return x;
            if (s > 0) {
                const long long lb = r ? fw.Kp : 1;
                j = (unsigned long long)((k - lb) / s) + 1;
            } else {
                const long long a = -s;
                long long room = (L - 1 - (long long)x) / a;
                if (!r && fw.Kp != KP_NONE) {
                    const long long q = (fw.Kp - 1 - k) / a + 1;
                    if (q < room) room = q;
                }
                j = room > 0 ? (unsigned long long)room : 0;
            }
        }
        if (j == 0) {
            const int y = fine_next(x, fw.d, fw.len);
            if (y == x) return x;
            x = y;
            --n;
            continue;
        }
        if (j > n) j = n;
        x = int((long long)x - (long long)j * s);
        n -= j;
    }
    return x;
}

// The same region walk in 32-bit arithmetic (64-bit division is a ~500-cycle software routine on the
// device; a receiver CTA walks the index a few windows at a time, on its critical path).  Valid when
// |D|, len and n are below 2^28 (fine_walk32_ok); non-integer steps only -- the integer case and anything
// out of range take fine_advance_dev.  Same decisions as fine_advance_dev, so the same exact result.
struct FineWalk32 {
    int len;
    float d;
    int D, Kp;      // Kp == 0x7fffffff: none
    int ok;
};
__device__ __forceinline__ FineWalk32 fine_walk32_make(const FineWalkDev &fw)
{
    FineWalk32 w;
    w.len = fw.len; w.d = fw.d;
    const long long lim = 1ll << 28;
    const bool kp_none = fw.Kp == 0x7fffffffffffffffll;
    w.ok = (!fw.integer && fw.D > -lim && fw.D < lim && fw.len < int(lim) && (kp_none || fw.Kp < lim)) ? 1 : 0;
    w.D = int(fw.D);
    w.Kp = kp_none ? 0x7fffffff : int(fw.Kp);
    return w;
}
__device__ __forceinline__ int fine_advance32(const FineWalk32 &fw, int x, unsigned n)
{
    const int L = fw.len;
    while (n > 0) {
        const int k = x - fw.D;
        unsigned j = 0;
        int s = 0;
        if (k >= 1) {
            const bool r = (k >= fw.Kp);
            s = fw.D + 1 - (r ? 1 : 0);
            if (s == 0) return x;
            if (s > 0) {
                const int lb = r ? fw.Kp : 1;
                j = unsigned(k - lb) / unsigned(s) + 1u;
            } else {
                const unsigned a = unsigned(-s);
                unsigned room = unsigned(L - 1 - x) / a;
                if (!r && fw.Kp != 0x7fffffff) {
                    const unsigned q = unsigned(fw.Kp - 1 - k) / a + 1u;
                    if (q < room) room = q;
                }
                j = room;
            }
        }
        if (j == 0) {
            const int y = fine_next(x, fw.d, L);
            if (y == x) return x;
            x = y;
            --n;
            continue;
        }
        if (j > n) j = n;
        x = x - int(j) * s;
        n -= j;
    }
    return x;
}

__global__ void fine_small_kernel(int32_t *__restrict__ X, const FineWalkDev fw, int x0, unsigned long long q_end)
{
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < q_end) X[q] = fine_advance_dev(fw, x0, q * kFineChunk);
}

} // namespace lorab200
