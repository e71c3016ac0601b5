// demod_fused.cuh -- the fast path: fused dechirp + 2^SF-point FFT + peak detect,
// persistent CTAs, TMA-staged input, FFT in registers with padded shared-memory
// exchanges, nothing but the 16/32-byte result record goes back to HBM.
//
// Replaces LoRaDemod.cpp:157-172 (dechirp loop + LoRaDetector::detect, i.e.
// kissfft.hh:77-157 + LoRaDetector.hpp:33-61) for dense batches of windows.
//
// Structure (tools/layout_model.py is the executable model of the index maps and
// proves the shared-memory accesses bank-conflict free):
//
//   * A window of N = 2^SF samples is owned by G = N/16 threads, 16 points each.
//     A "group" is max(32, G) threads and owns a private ring of TMA stages; one
//     tile = GROUP*16 samples (4 / 2 / 1 windows for SF 7 / 8 / >=9).  Groups never
//     synchronise with each other: SF<=9 groups are single warps (__syncwarp only),
//     larger ones use their own named barrier.
//   * Lane 0 of a group issues `cp.async.bulk` (TMA) global->shared for tile k+S as
//     soon as the group has finished tile k; consumers wait on the stage's mbarrier.
//   * kissfft's stage list (radix 4,...,4[,2], deepest first) is cut into passes of
//     <= 4 bits that run entirely in registers: pass 0 covers the b0 lowest bits of
//     the output index (no thread-dependent twiddles), later passes are radix-16
//     (two radix-4 stages, 15 per-thread twiddles kept in registers for the whole
//     persistent loop).  Between passes values go through shared memory in a padded
//     natural-order layout so that both the scatter and the gather are conflict free
//     and every address is "thread base + immediate".
//   * Exact mode: every float op is an explicitly rounded mul/add/sub in the order of
//     the reference's operand graph, twiddles and chirp come from the host-built
//     tables => FFT bins are bit-identical to kissfft (except the sign of zeros:
//     the multiplications by tw[0] = 1+0j are skipped).
#pragma once
#include "common.cuh"

namespace lorab200 {

// ---- compile-time geometry ---------------------------------------------------------
template <int SF>
struct FusedCfg {
    static constexpr int N = 1 << SF;
    static constexpr int NP = (SF + 3) / 4;             // passes
    static constexpr int B0 = SF - 4 * (NP - 1);        // bits handled by pass 0
    static constexpr int R0 = 1 << B0;                  // points per pass-0 sub-FFT
    static constexpr int E = 16 / R0;                   // pass-0 sub-FFTs per thread
    static constexpr int G = N / 16;                    // threads per window
    static constexpr int GROUP = G < 32 ? 32 : G;       // threads per group (tile owner)
    static constexpr int WPT = GROUP / G;               // windows per tile
    static constexpr int TILE = GROUP * 16;             // samples per tile
    static constexpr int CTA = 256;
    static constexpr int GROUPS = CTA / GROUP;          // groups per CTA
    static constexpr bool HAS_X = (NP == 3);            // separate exchange buffer for exchange 0
    static constexpr bool CHIRP_REGS = (NP == 2);       // chirp table slice kept in registers
    static constexpr int STAGES = (SF == 7 || HAS_X) ? 2 : 3;

    // exchange layout: natural output index j -> padded element offset inside the window
    __host__ __device__ static constexpr int pad(int j)
    {
        return SF == 7 ? j + (j >> 3) + 2 * (j >> 5) : j + (j >> 4) + (j >> 8);
    }
    static constexpr int WSTRIDE = SF == 7 ? 152 : pad(N - 1) + 1; // elements per window (exchange)
    static constexpr int XELEMS = WPT * WSTRIDE;                    // exchange footprint of a tile
    static constexpr int STAGE_ELEMS = ((XELEMS > TILE ? XELEMS : TILE) + 15) / 16 * 16;
    static constexpr int STAGE_BYTES = STAGE_ELEMS * 8;
    static constexpr int GROUP_BYTES = (STAGES + (HAS_X ? 1 : 0)) * STAGE_BYTES;
    static constexpr int BAR_BYTES = 128 * ((GROUPS * STAGES * 8 + 127) / 128);
    static constexpr int RED_BYTES = GROUPS * 8 * 32;   // cross-warp reduction scratch (GROUP > 32)
    static constexpr int SMEM_BYTES = BAR_BYTES + GROUPS * GROUP_BYTES + RED_BYTES;
};

// Uniform twiddles used inside pass 0 (kissfft.hh:141-143 with m = 2 or 4): passed by
// value in the kernel parameters so they are constant-bank operands.
struct Pass0Tw {
    float2 w[3][3]; // w[q-1][k-1] = tw[q * k * N / R0'], see fill_pass0_tw
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
template <int GROUP>
__device__ __forceinline__ void group_sync(int group_in_cta)
{
    if constexpr (GROUP == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group_in_cta + 1), "n"(GROUP) : "memory");
}

__device__ __forceinline__ float2 lds64(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds128(uint32_t addr, float2 &a, float2 &b)
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(b.x), "=f"(b.y) : "r"(addr));
}
__device__ __forceinline__ void sts64(uint32_t addr, float2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// base-4 digit reversal of v over `bits` bits (bits even): kissfft's input permutation
// restricted to the radix-4 stages (kissfft.hh:90-94,101-103).
__host__ __device__ constexpr int digit_rev4(int v, int bits)
{
    int r = 0;
    for (int b = 0; b < bits; b += 2) { r = (r << 2) | (v & 3); v >>= 2; }
    return r;
}

// pass-0 register slot c (input order, stride N/R0) -> low output-index bits x
template <int B0>
__host__ __device__ constexpr int slot_to_x(int c)
{
    return B0 <= 2 ? c : (B0 == 3 ? (((c & 3) << 1) | (c >> 2)) : (((c & 3) << 2) | (c >> 2)));
}

// Two radix-4 stages on 16 register values r[y], y = (qd << 2) | qc, with the
// per-thread twiddles twa[qc-1] (first stage, same for the 4 butterflies) and
// twb[qc][qd-1] (second stage).  kissfft.hh:137-157.
__device__ __forceinline__ void radix16(float2 (&r)[16], const float2 (&twa)[3], const float2 (&twb)[4][3])
{
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
        float2 a1 = cx_mul(r[qd * 4 + 1], twa[0]);
        float2 a2 = cx_mul(r[qd * 4 + 2], twa[1]);
        float2 a3 = cx_mul(r[qd * 4 + 3], twa[2]);
        bfly4(r[qd * 4 + 0], a1, a2, a3);
        r[qd * 4 + 1] = a1; r[qd * 4 + 2] = a2; r[qd * 4 + 3] = a3;
    }
#pragma unroll
    for (int qc = 0; qc < 4; ++qc) {
        float2 a1 = cx_mul(r[4 + qc], twb[qc][0]);
        float2 a2 = cx_mul(r[8 + qc], twb[qc][1]);
        float2 a3 = cx_mul(r[12 + qc], twb[qc][2]);
        bfly4(r[qc], a1, a2, a3);
        r[4 + qc] = a1; r[8 + qc] = a2; r[12 + qc] = a3;
    }
}

// Load the 15 twiddles of a radix-16 pass whose first stage has m = 2^A:
//   first stage : tw[qc * lo * N/(4m)]              (k = lo)
//   second stage: tw[qd * (lo + qc*m) * N/(16m)]    (k = lo + qc*m)
template <int SF, int A>
__device__ __forceinline__ void load_pass_tw(const float2 *__restrict__ tw, int lo, float2 (&twa)[3], float2 (&twb)[4][3])
{
    constexpr int N = 1 << SF;
    constexpr int M = 1 << A;
    constexpr int FSA = N / (4 * M);
    constexpr int FSB = N / (16 * M);
#pragma unroll
    for (int q = 1; q < 4; ++q) twa[q - 1] = tw[q * lo * FSA];
#pragma unroll
    for (int qc = 0; qc < 4; ++qc)
#pragma unroll
        for (int qd = 1; qd < 4; ++qd) twb[qc][qd - 1] = tw[qd * (lo + qc * M) * FSB];
}

// pick r[slot] with a compile-time-indexed switch (keeps r[] in registers)
__device__ __forceinline__ float2 pick16(const float2 (&r)[16], int slot)
{
    float2 v = r[0];
    switch (slot) {
    case 1: v = r[1]; break;   case 2: v = r[2]; break;   case 3: v = r[3]; break;
    case 4: v = r[4]; break;   case 5: v = r[5]; break;   case 6: v = r[6]; break;
    case 7: v = r[7]; break;   case 8: v = r[8]; break;   case 9: v = r[9]; break;
    case 10: v = r[10]; break; case 11: v = r[11]; break; case 12: v = r[12]; break;
    case 13: v = r[13]; break; case 14: v = r[14]; break; case 15: v = r[15]; break;
    default: break;
    }
    return v;
}

template <int SF>
__global__ void __launch_bounds__(256, 2)
demod_fused_kernel(const DemodParams p, const Pass0Tw p0tw)
{
    using C = FusedCfg<SF>;
    constexpr int N = C::N, B0 = C::B0, R0 = C::R0, E = C::E, G = C::G, GROUP = C::GROUP;
    constexpr int WPT = C::WPT, TILE = C::TILE, STAGES = C::STAGES;

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem_base = smem_u32(smem);

    const int tid = threadIdx.x;
    const int grp = tid / GROUP;              // group within the CTA
    const int tg = tid % GROUP;               // thread within the group
    const int wsub = tg / G;                  // window within the tile
    const int t = tg % G;                     // thread within the window
    const int lane = tid & 31;

    const uint32_t bar0 = smem_base + uint32_t(grp * STAGES) * 8u;
    const uint32_t gbase = smem_base + C::BAR_BYTES + uint32_t(grp) * C::GROUP_BYTES;
    const uint32_t xbase = gbase + STAGES * C::STAGE_BYTES; // only if HAS_X
    double *red_sum = reinterpret_cast<double *>(smem + C::BAR_BYTES + C::GROUPS * C::GROUP_BYTES) + grp * 32;
    // the same scratch viewed as (float value, int index) pairs after the sums
    // (layout per group: 8 doubles | 8 floats | 8 ints | neighbours)
    float *red_val = reinterpret_cast<float *>(red_sum + 8);
    int *red_idx = reinterpret_cast<int *>(red_val + 8);
    float2 *red_nb = reinterpret_cast<float2 *>(red_idx + 8); // [0]=left [1]=right

    const unsigned long long n_tiles = (p.n_windows + WPT - 1) / WPT;
    const unsigned long long gid = (unsigned long long)blockIdx.x * C::GROUPS + grp;
    const unsigned long long gstride = (unsigned long long)gridDim.x * C::GROUPS;

    // ---- one-time setup ------------------------------------------------------------
    if (tg == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + s * 8, 1);
        fence_mbar_init();
    }
    group_sync<GROUP>(grp);

    auto issue = [&](unsigned long long tile, int s) {
        // lane 0 of the group: TMA the whole tile (contiguous: stride == N)
        const unsigned long long w0 = tile * WPT;
        unsigned long long nw = p.n_windows - w0;
        if (nw > WPT) nw = WPT;
        const uint32_t bytes = uint32_t(nw) * N * 8u;
        mbar_expect_tx(bar0 + s * 8, bytes);
        tma_load_1d(gbase + s * C::STAGE_BYTES, p.iq + w0 * N, bytes, bar0 + s * 8);
    };
    if (tg == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const unsigned long long tile = gid + (unsigned long long)s * gstride;
            if (tile < n_tiles) issue(tile, s);
        }
    }

    // per-thread pass-0 input offsets (elements, linear stage layout) and chirp
    // i(e, c) = ilow(e) + (N/R0)*c,  ilow = t (E==1) or ehi*2G + 2t + e0
    const int ilow0 = (E == 1) ? t : 2 * t;
    float2 chirp_r[C::CHIRP_REGS ? 16 : 1];
    if constexpr (C::CHIRP_REGS) {
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int c = 0; c < R0; ++c) {
                const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
                chirp_r[e * R0 + c] = p.chirp[il + (N / R0) * c];
            }
    }
    // twiddles of the radix-16 passes (registers for the whole persistent loop)
    float2 tw1a[3], tw1b[4][3];
    float2 tw2a[3], tw2b[4][3];
    // pass-1 thread coordinates: lo = low B0 bits of the output index, hi = the rest
    // NP==2: t1 = lo (G == 2^B0).  NP==3: t1 = lo | hi << B0.
    const int lo1 = t & (R0 - 1);
    const int hi1 = t >> B0;
    load_pass_tw<SF, B0>(p.tw, lo1, tw1a, tw1b);
    if constexpr (C::NP == 3) load_pass_tw<SF, B0 + 4>(p.tw, t, tw2a, tw2b);

    // exchange addresses (bytes, relative to a window's exchange base)
    // pass-0 scatter: j = x(c) | digit_rev(ilow(e)) << B0
    uint32_t sc0[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
        sc0[e] = uint32_t(C::pad(digit_rev4(il, SF - B0) << B0)) * 8u;
    }
    // pass-1 gather/scatter: j = lo | y << B0 | hi << (B0+4)
    const uint32_t ga1 = uint32_t(C::pad(lo1 | (hi1 << (B0 + 4)))) * 8u;
    // pass-2 gather: j = t | y << (B0+4)
    const uint32_t ga2 = uint32_t(C::pad(t)) * 8u;
    const uint32_t wx = uint32_t(wsub) * C::WSTRIDE * 8u; // window offset inside exchange layout

    uint32_t phase = 0; // bit s = parity to wait for on stage s
    int s = 0;
    for (unsigned long long tile = gid; tile < n_tiles; tile += gstride) {
        const uint32_t stage = gbase + uint32_t(s) * C::STAGE_BYTES;
        const unsigned long long w = tile * WPT + wsub;     // this thread's window
        const bool live = w < p.n_windows;

        // chirp for this tile via the read-only path when it is not register resident
        float2 r[16];
        float2 ch[C::CHIRP_REGS ? 1 : 16];
        if constexpr (!C::CHIRP_REGS) {
#pragma unroll
            for (int e = 0; e < E; ++e)
#pragma unroll
                for (int c = 0; c < R0; ++c) {
                    const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
                    ch[e * R0 + c] = __ldg(p.chirp + il + (N / R0) * c);
                }
        }

        mbar_wait(bar0 + s * 8, (phase >> s) & 1u);
        phase ^= 1u << s;

        // ---- pass 0: load (linear layout), dechirp ---------------------------------------
        {
            const uint32_t in0 = stage + (uint32_t(wsub) * N + uint32_t(ilow0)) * 8u;
            if constexpr (E == 1) {
#pragma unroll
                for (int c = 0; c < R0; ++c) r[c] = lds64(in0 + uint32_t((N / R0) * c) * 8u);
            } else {
#pragma unroll
                for (int eh = 0; eh < E / 2; ++eh)
#pragma unroll
                    for (int c = 0; c < R0; ++c)
                        lds128(in0 + uint32_t(eh * 2 * G + (N / R0) * c) * 8u,
                               r[(2 * eh) * R0 + c], r[(2 * eh + 1) * R0 + c]);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if constexpr (C::CHIRP_REGS) r[k] = cx_mul(r[k], chirp_r[k]);
                else r[k] = cx_mul(r[k], ch[k]);
            }
            if (p.dec && live) {
#pragma unroll
                for (int e = 0; e < E; ++e)
#pragma unroll
                    for (int c = 0; c < R0; ++c) {
                        const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
                        p.dec[w * N + il + (N / R0) * c] = r[e * R0 + c];
                    }
            }
        }
        // ---- pass 0 butterflies (deepest kissfft stages; tw[0] multiplies skipped) ---------
#pragma unroll
        for (int e = 0; e < E; ++e) {
            float2 *v = r + e * R0;
            if constexpr (B0 == 1) {
                bfly2(v[0], v[1]);
            } else if constexpr (B0 == 2) {
                bfly4(v[0], v[1], v[2], v[3]);
            } else if constexpr (B0 == 3) {
                // radix-2 (m=1) on pairs (c, c+4), then radix-4 (m=2) over c = q + 4k
#pragma unroll
                for (int a = 0; a < 4; ++a) bfly2(v[a], v[a + 4]);
                bfly4(v[0], v[1], v[2], v[3]);                       // k = 0: tw[0]
                float2 a1 = cx_mul(v[5], p0tw.w[0][0]);
                float2 a2 = cx_mul(v[6], p0tw.w[1][0]);
                float2 a3 = cx_mul(v[7], p0tw.w[2][0]);
                bfly4(v[4], a1, a2, a3);
                v[5] = a1; v[6] = a2; v[7] = a3;
            } else {
                // radix-4 (m=1) over qa for each qb (slots qb + 4*qa), then radix-4 (m=4)
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) bfly4(v[qb], v[qb + 4], v[qb + 8], v[qb + 12]);
                bfly4(v[0], v[1], v[2], v[3]);                       // k = 0
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                    float2 a1 = cx_mul(v[4 * k + 1], p0tw.w[0][k - 1]);
                    float2 a2 = cx_mul(v[4 * k + 2], p0tw.w[1][k - 1]);
                    float2 a3 = cx_mul(v[4 * k + 3], p0tw.w[2][k - 1]);
                    bfly4(v[4 * k], a1, a2, a3);
                    v[4 * k + 1] = a1; v[4 * k + 2] = a2; v[4 * k + 3] = a3;
                }
            }
        }

        // ---- exchange 0 -------------------------------------------------------------------
        const uint32_t x0 = (C::HAS_X ? xbase : stage) + wx;
        if constexpr (!C::HAS_X) group_sync<GROUP>(grp); // in place: everyone has read the linear tile
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int c = 0; c < R0; ++c)
                sts64(x0 + sc0[e] + uint32_t(C::pad(slot_to_x<B0>(c))) * 8u, r[e * R0 + c]);
        group_sync<GROUP>(grp);
#pragma unroll
        for (int y = 0; y < 16; ++y) r[y] = lds64(x0 + ga1 + uint32_t(C::pad(y << B0)) * 8u);

        // ---- pass 1 (and exchange 1 + pass 2) ---------------------------------------------------
        radix16(r, tw1a, tw1b);
        if constexpr (C::NP == 3) {
            const uint32_t x1 = stage + wx; // in place in the stage: all its loads completed before the sync above
#pragma unroll
            for (int y = 0; y < 16; ++y) sts64(x1 + ga1 + uint32_t(C::pad(y << B0)) * 8u, r[y]);
            group_sync<GROUP>(grp);
#pragma unroll
            for (int y = 0; y < 16; ++y) r[y] = lds64(x1 + ga2 + uint32_t(C::pad(y << (B0 + 4))) * 8u);
            radix16(r, tw2a, tw2b);
        }

        // ---- peak scan: thread holds bins j = t + G*y, y ascending ----------------------------------
        float bv = 0.0f; int by = 0;
        double sum0 = 0.0, sum1 = 0.0;
#pragma unroll
        for (int y = 0; y < 16; ++y) {
            const float m2 = cx_mag2(r[y]);
            if (y & 1) sum1 += double(m2); else sum0 += double(m2);
            if (m2 > bv) { bv = m2; by = y; }
        }
        double sum = sum0 + sum1;
        int bi = t + G * by;
        if (p.fft && live) {
#pragma unroll
            for (int y = 0; y < 16; ++y) p.fft[w * N + t + G * y] = r[y];
        }
        // reduce over the window's G threads
        constexpr int WIDTH = G < 32 ? G : 32;
#pragma unroll
        for (int o = WIDTH / 2; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            argmax_merge(bv, bi, ov, oi);
        }
        if constexpr (G > 32) {
            // cross-warp: warp leaders publish, every thread folds the G/32 partials
            const int wig = tg >> 5; // warp in group
            if (lane == 0) { red_sum[wig] = sum; red_val[wig] = bv; red_idx[wig] = bi; }
            group_sync<GROUP>(grp);
            sum = red_sum[0]; bv = red_val[0]; bi = red_idx[0];
#pragma unroll
            for (int k = 1; k < G / 32; ++k) { sum += red_sum[k]; argmax_merge(bv, bi, red_val[k], red_idx[k]); }
        }
        // neighbours of the peak: bin (bi-1) mod N and (bi+1) mod N live in other threads' registers
        const int jl = bi > 0 ? bi - 1 : N - 1, jr = bi < N - 1 ? bi + 1 : 0;
        const int tl = jl & (G - 1), yl = jl / G, tr = jr & (G - 1), yr = jr / G;
        float2 nbv = make_float2(0.f, 0.f);
        if (t == tl) nbv = pick16(r, yl);
        else if (t == tr) nbv = pick16(r, yr);
        float2 left, right;
        if constexpr (G <= 32) {
            const int base = lane - t; // first lane of this window
            left.x = __shfl_sync(0xffffffffu, nbv.x, base + tl);
            left.y = __shfl_sync(0xffffffffu, nbv.y, base + tl);
            right.x = __shfl_sync(0xffffffffu, nbv.x, base + tr);
            right.y = __shfl_sync(0xffffffffu, nbv.y, base + tr);
        } else {
            if (t == tl) red_nb[0] = nbv;
            else if (t == tr) red_nb[1] = nbv;
            group_sync<GROUP>(grp);
            left = red_nb[0]; right = red_nb[1];
        }
        if (t == 0 && live) {
            RawResult rr;
            rr.index = bi; rr.max_mag2 = bv; rr.total = sum; rr.left = left; rr.right = right;
            if (p.raw) p.raw[w] = rr;
            if (p.out) finalize_device(rr, p.power_scale, &p.out[w]);
        }

        // ---- stage s is free: refill it with the tile S iterations ahead -------------------------
        group_sync<GROUP>(grp);
        if (tg == 0) {
            const unsigned long long nt = tile + (unsigned long long)STAGES * gstride;
            if (nt < n_tiles) { fence_proxy_async(); issue(nt, s); }
        }
        s = (s + 1 == STAGES) ? 0 : s + 1;
    }
}

// ---- host side -------------------------------------------------------------------------
inline bool fused_applicable(int sf, const DemodParams &p)
{
    if (sf < 7 || sf > 12) return false;
    const unsigned long long n = 1ull << sf;
    if (p.stride != n && p.n_windows > 1) return false;              // dense batches only
    if ((reinterpret_cast<uintptr_t>(p.iq) & 15u) != 0) return false; // TMA needs 16-byte alignment
    if (p.fine != nullptr) return false;                              // fine-tune path: simple kernel
    return true;
}

inline void fill_pass0_tw(int sf, const float2 *tw_host, Pass0Tw &o)
{
    // B0 == 3: second pass-0 stage has m = 2, fstride = N/8, k = 1:      w[q-1][0] = tw[q*N/8]
    // B0 == 4: second pass-0 stage has m = 4, fstride = N/16, k = 1..3:  w[q-1][k-1] = tw[q*k*N/16]
    const int n = 1 << sf;
    const int np = (sf + 3) / 4, b0 = sf - 4 * (np - 1);
    for (int q = 1; q < 4; ++q)
        for (int k = 1; k < 4; ++k) {
            float2 v = make_float2(1.f, 0.f);
            if (b0 == 3 && k == 1) v = tw_host[q * n / 8];
            if (b0 == 4) v = tw_host[q * k * n / 16];
            o.w[q - 1][k - 1] = v;
        }
}

template <int SF>
cudaError_t launch_fused_sf(const DemodParams &p, const Pass0Tw &tw0, cudaStream_t s, int sm_count)
{
    using C = FusedCfg<SF>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(demod_fused_kernel<SF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const unsigned long long n_tiles = (p.n_windows + C::WPT - 1) / C::WPT;
    unsigned long long ctas = (n_tiles + C::GROUPS - 1) / C::GROUPS;
    const unsigned long long cap = 2ull * (unsigned long long)sm_count; // persistent: 2 CTAs per SM
    if (ctas > cap) ctas = cap;
    demod_fused_kernel<SF><<<unsigned(ctas), C::CTA, C::SMEM_BYTES, s>>>(p, tw0);
    return cudaGetLastError();
}

inline cudaError_t launch_fused(int sf, const DemodParams &p, const Pass0Tw &tw0, cudaStream_t s, int sm_count)
{
    switch (sf) {
    case 7: return launch_fused_sf<7>(p, tw0, s, sm_count);
    case 8: return launch_fused_sf<8>(p, tw0, s, sm_count);
    case 9: return launch_fused_sf<9>(p, tw0, s, sm_count);
    case 10: return launch_fused_sf<10>(p, tw0, s, sm_count);
    case 11: return launch_fused_sf<11>(p, tw0, s, sm_count);
    case 12: return launch_fused_sf<12>(p, tw0, s, sm_count);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace lorab200
