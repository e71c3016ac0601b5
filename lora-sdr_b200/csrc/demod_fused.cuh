// demod_fused.cuh -- the fast path: fused dechirp + 2^SF-point FFT + peak detect,
// persistent CTAs, TMA-staged input, FFT in registers on the packed FP32 pipe with
// padded shared-memory exchanges; nothing but the per-window result record goes
// back to HBM.
//
// Replaces LoRaDemod.cpp:157-172 (dechirp loop + LoRaDetector::detect, i.e.
// kissfft.hh:77-157 + LoRaDetector.hpp:33-61) for dense batches of windows.
//
// Structure (tools/layout_model.py is the executable model of the index maps and
// proves the shared-memory accesses bank-conflict free):
//
//   * A window of N = 2^SF samples is owned by G = N/16 threads, 16 points each.
//     A "group" is max(32, G) threads; it owns a private ring of linear TMA stages,
//     one padded exchange buffer and a contiguous range of tiles (one tile =
//     GROUP*16 samples = 4 / 2 / 1 windows for SF 7 / 8 / >=9).  Groups never
//     synchronise with each other: SF<=9 groups are single warps (__syncwarp only),
//     larger ones use their own named barrier.
//   * Thread 0 of a group issues `cp.async.bulk` (TMA, SASS UBLKCP) global->shared
//     into a stage as soon as the group has copied that stage into registers;
//     consumers wait on the stage's mbarrier.
//   * kissfft's stage list (radix 4,...,4[,2], deepest first) is cut into passes of
//     <= 4 bits that run entirely in registers: pass 0 covers the b0 lowest bits of
//     the output index (no thread-dependent twiddles), later passes are radix-16
//     (two radix-4 stages, 15 per-thread twiddles kept in registers for the whole
//     persistent loop).  Between passes values go through the exchange buffer in a
//     padded natural-order layout so that both the scatter and the gather are
//     conflict free and every address is "thread base + immediate".
//   * Exact mode on the packed pipe (packed.cuh): each complex value is an aligned
//     register pair, add/sub are one FADD2, a complex multiply is FMUL2+FMUL2+FFMA2,
//     multiplication by -j is an operand modifier; every rounding is the one the
//     reference's std::complex<float> code performs, twiddles and chirp come from
//     the host-built tables => FFT bins are bit-identical to kissfft (except the
//     sign of zeros: the multiplications by tw[0] = 1+0j are skipped).
//   * Raw detector outputs are parked in shared memory and finalised 32 windows at
//     a time by a full warp (one coalesced 512-byte store) instead of one lane per
//     window running the double-precision log10 alone.
#pragma once
#include "common.cuh"
#include "fine_orbit.h"
#include "packed.cuh"

#ifndef LORAB200_ABL
#define LORAB200_ABL 0   // timing ablations (tools/build_variant.sh); results are wrong when non-zero
#endif

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
    // Occupancy plan: 2 CTAs per SM (<= 128 registers), two TMA stages per group.
    // (Measured alternative for the two-pass SFs, -DLORAB200_TRY_3CTA: 3 CTAs per SM at
    // 80 registers with one stage per group is slower -- SF7 3083 vs 3674 Msym/s.)
#ifdef LORAB200_TRY_3CTA
    static constexpr int CTAS_PER_SM = (NP == 2) ? 3 : 2;
#else
    static constexpr int CTAS_PER_SM = 2;
#endif
    // SF >= 11: one stage (a window takes far longer to process than a tile takes to arrive,
    // and the stage is handed back to the TMA right after it has been copied to registers)
    static constexpr int STAGES = (CTAS_PER_SM == 3 || SF >= 11) ? 1 : 2;
    static constexpr bool CHIRP_SMEM = true;            // chirp table staged in shared memory (TMA)
    // raw results parked per group before a flush (a multiple of WPT)
    static constexpr int RAWCAP = (CTAS_PER_SM == 3) ? (SF == 7 ? 12 : 24) : 32;

    // exchange layout: natural output index j -> padded element offset inside the window
    __host__ __device__ static constexpr int pad(int j)
    {
        return SF == 7 ? j + (j >> 3) + 2 * (j >> 5) : j + (j >> 4) + (j >> 8);
    }
    static constexpr int WSTRIDE = SF == 7 ? 152 : pad(N - 1) + 1; // elements per window (exchange)
    static constexpr int STAGE_BYTES = TILE * 8;
    static constexpr int X_BYTES = (WPT * WSTRIDE * 8 + 127) / 128 * 128;
    static constexpr int RAW_BYTES = RAWCAP * 32;
    static constexpr int RED_BYTES = GROUP > 32 ? 256 : 0; // cross-warp reduction scratch
    static constexpr int GROUP_BYTES = STAGES * STAGE_BYTES + X_BYTES + RAW_BYTES + RED_BYTES;
    static constexpr int BAR_BYTES = 16 * (((GROUPS * STAGES + 1) * 8 + 15) / 16); // stage barriers + chirp barrier
    static constexpr int CHIRP_BYTES = CHIRP_SMEM ? N * 8 : 0;
    static constexpr int SMEM_BYTES = BAR_BYTES + CHIRP_BYTES + GROUPS * GROUP_BYTES;
    // the planned number of CTAs must fit in one SM: 228 KB per SM, 1 KB reserved per CTA
    static_assert(CTAS_PER_SM * (SMEM_BYTES + 1024) <= 233472, "shared memory budget");
    static_assert(RAWCAP % WPT == 0 && RAWCAP <= 32, "RAWCAP");
    // FINE_WALK parks N fine-tune indices per window in the exchange buffer while it is idle:
    // 32-bit words at fine_pad(j); window stride chosen so two windows of a warp use disjoint banks
    static constexpr int FINE_WSTRIDE = N + N / 8 + (SF == 8 ? 16 : 0); // words
    static_assert(WPT * FINE_WSTRIDE * 4 <= X_BYTES && (FINE_WSTRIDE * 4) % 16 == 0, "fine index area");
};

// Uniform twiddles used inside pass 0 (kissfft.hh:141-143 with m = 2 or 4): passed by
// value in the kernel parameters (constant bank / uniform registers).
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
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier.
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
// programmatic dependent launch: let the next kernel in the stream start its prologue
// while this grid drains / wait for the previous grid before touching its memory
__device__ __forceinline__ void pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_prior_grid()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <int GROUP>
__device__ __forceinline__ void group_sync(int group_in_cta)
{
    if constexpr (GROUP == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group_in_cta + 1), "n"(GROUP) : "memory");
}

__device__ __forceinline__ c64 lds64(uint32_t addr)
{
    c64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds128(uint32_t addr, c64 &a, c64 &b)
{
    float x, y, z, w;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr));
    a = cpack(x, y);
    b = cpack(z, w);
}
__device__ __forceinline__ void sts64(uint32_t addr, c64 v)
{
    const float2 f = cunpack(v);
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(f.x), "f"(f.y) : "memory");
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

// pass-0 scatter offset (elements) of sub-FFT e relative to sub-FFT 0 of the same thread
template <int SF>
__host__ __device__ constexpr int sc0_eoff(int e)
{
    using C = FusedCfg<SF>;
    return C::E == 1 ? 0 : C::pad(digit_rev4((e >> 1) * 2 * C::G + (e & 1), SF - C::B0) << C::B0);
}
template <int SF>
__host__ __device__ constexpr bool sc0_split_ok()
{
    using C = FusedCfg<SF>;
    if constexpr (C::E > 1)
    for (int t = 0; t < C::G; ++t)
        for (int e = 0; e < C::E; ++e)
            for (int c = 0; c < C::R0; ++c) {
                const int il = (e >> 1) * 2 * C::G + 2 * t + (e & 1);
                const int full = C::pad((digit_rev4(il, SF - C::B0) << C::B0) | slot_to_x<C::B0>(c));
                const int split = C::pad(digit_rev4(2 * t, SF - C::B0) << C::B0) + sc0_eoff<SF>(e) + C::pad(slot_to_x<C::B0>(c));
                if (full != split) return false;
            }
    return true;
}

// Two radix-4 stages on 16 register values r[y], y = (qd << 2) | qc, with the
// per-thread twiddles twa[qc-1] (first stage, same for the 4 butterflies) and
// twb[qc][qd-1] (second stage).  kissfft.hh:137-157.  Per 16 points: 48 FMUL2 +
// 24 FFMA2 on the FMA pipe, 64 FADD2 on the add pipe.
__device__ __forceinline__ void radix16(c64 (&r)[16], const float2 (&twa)[3], const float2 (&twb)[4][3])
{
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
        r[qd * 4 + 1] = pcx_mul(r[qd * 4 + 1], twa[0]);
        r[qd * 4 + 2] = pcx_mul(r[qd * 4 + 2], twa[1]);
        r[qd * 4 + 3] = pcx_mul(r[qd * 4 + 3], twa[2]);
        p_bfly4(r[qd * 4 + 0], r[qd * 4 + 1], r[qd * 4 + 2], r[qd * 4 + 3]);
    }
#pragma unroll
    for (int qc = 0; qc < 4; ++qc) {
        r[4 + qc] = pcx_mul(r[4 + qc], twb[qc][0]);
        r[8 + qc] = pcx_mul(r[8 + qc], twb[qc][1]);
        r[12 + qc] = pcx_mul(r[12 + qc], twb[qc][2]);
        p_bfly4(r[qc], r[4 + qc], r[8 + qc], r[12 + qc]);
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

// r[slot] for a run-time slot, as a binary select tree (keeps r[] in registers, no divergence)
__device__ __forceinline__ c64 pick16(const c64 (&r)[16], int slot)
{
    c64 a[8], b[4], c[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (slot & 1) ? r[2 * i + 1] : r[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (slot & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = (slot & 4) ? b[2 * i + 1] : b[2 * i];
    return (slot & 8) ? c[1] : c[0];
}

// DBG : also write the fft / dec debug dumps (p.fft, p.dec may be null individually)
// FINE: fine-tune multiply of LoRaDemod.cpp:159
//   FINE_NONE  no fine-tune table (the detector-only form, TestDetector.cpp)
//   FINE_CONST one constant phasor for the whole batch (_finefreqError == 0: the index never moves)
//   FINE_WALK  the index walks once per sample (LoRaDemod.cpp:160-162).  The state at the start
//              of every 16-sample chunk comes from the pointer-jumping prefix (fine_prefix.cuh);
//              thread t of a window walks chunk t with the literal recurrence and parks the 16
//              indices in the (idle) exchange buffer, from where every thread picks up the
//              indices of its own pass-0 samples; the phasors are gathered from the table in L2.
//   FINE_ORBIT the same walking index, evaluated in O(1) per sample from the orbit description of
//              fine_orbit.h (a kernel parameter): no prefix kernels, no index parking, no extra
//              barriers, and the programmatic-launch wait can move down to the first flush.  Used
//              whenever the step has an orbit (all but steps within 2^-5 of an integer); FINE_WALK is
//              the fallback.
enum { FINE_NONE = 0, FINE_CONST = 1, FINE_WALK = 2, FINE_ORBIT = 3 };

// the orbit travels as a kernel parameter only in the instantiations that use it
template <int FINE> struct OrbitArg { };
template <> struct OrbitArg<FINE_ORBIT> { FineOrbit o; };

// word offset of fine index j inside a window's exchange area: +4 words every 32 so that the
// 64-byte per-thread chunks are written conflict free with 128-bit stores
__host__ __device__ constexpr int fine_pad(int j) { return j + 4 * (j >> 5); }
__device__ __forceinline__ int fine_pad_dyn(int j) { return j + 4 * (j >> 5); }

// MULTI: the batch is a list of segments (SegList, common.cuh); a tile never straddles two
//        segments (the host requires seg_windows % WPT == 0)
template <int SF, bool DBG, int FINE, bool MULTI = false>
__global__ void __launch_bounds__(256, FusedCfg<SF>::CTAS_PER_SM)
demod_fused_kernel(const DemodParams p, const Pass0Tw p0tw, const float2 fine_phasor, const SegList<MULTI> segs,
                   const OrbitArg<FINE> orb)
{
    constexpr bool FINE0 = (FINE == FINE_CONST);
    using C = FusedCfg<SF>;
    constexpr int N = C::N, B0 = C::B0, R0 = C::R0, E = C::E, G = C::G, GROUP = C::GROUP;
    constexpr int WPT = C::WPT, STAGES = C::STAGES;

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem_base = smem_u32(smem);

    const int tid = threadIdx.x;
    // group within the CTA, from the warp index read through a shuffle from lane 0: the compiler then
    // knows it is warp-uniform and keeps the group's stage / barrier addresses and all the tile and
    // TMA-source arithmetic in uniform registers (SF7: 118 -> 112 registers, 600 -> 570 instructions
    // per tile, no ELECT / R2UR loop around the bulk copy; +3.5 % SF7, +7 % SF8, +8 % SF9)
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int grp = GROUP <= 32 ? warp_u : warp_u / (GROUP / 32);
    const int tg = tid % GROUP;               // thread within the group
    const int wsub = tg / G;                  // window within the tile
    const int t = tg % G;                     // thread within the window
    const int lane = tid & 31;

    const uint32_t bar0 = smem_base + uint32_t(grp * STAGES) * 8u;
    const uint32_t chirp_s = smem_base + C::BAR_BYTES;
    unsigned char *gptr = smem + C::BAR_BYTES + C::CHIRP_BYTES + grp * C::GROUP_BYTES;
    const uint32_t gbase = smem_base + C::BAR_BYTES + C::CHIRP_BYTES + uint32_t(grp) * C::GROUP_BYTES;
    const uint32_t xbase = gbase + STAGES * C::STAGE_BYTES;
    RawResult *rawbuf = reinterpret_cast<RawResult *>(gptr + STAGES * C::STAGE_BYTES + C::X_BYTES);
    double *red_sum = reinterpret_cast<double *>(gptr + STAGES * C::STAGE_BYTES + C::X_BYTES + C::RAW_BYTES);
    // (red_sum: up to 8 cross-warp partials of 16 bytes: {double sum, float bits, int index})

    // ---- this group's contiguous range of tiles --------------------------------------------
    const uint32_t n_windows = uint32_t(p.n_windows);
    const uint32_t n_tiles = (n_windows + WPT - 1) / WPT;
    const uint32_t n_groups = gridDim.x * C::GROUPS;
    const uint32_t gid = blockIdx.x * C::GROUPS + grp;
    const uint32_t per = n_tiles / n_groups, rem = n_tiles % n_groups;
    const uint32_t tile0 = gid * per + (gid < rem ? gid : rem);
    const uint32_t my_tiles = per + (gid < rem ? 1u : 0u);

    // ---- one-time setup ---------------------------------------------------------------------
    // Everything up to pdl_wait_prior_grid() only touches the tables (constant for the life
    // of the context), so under programmatic dependent launch it overlaps the tail of the
    // previous demod kernel in the stream.
    pdl_launch_dependents();
    const uint32_t cbar = smem_base + uint32_t(C::GROUPS * STAGES) * 8u; // chirp-table barrier (CTA wide)

    auto issue = [&](uint32_t k, int s) {
        // TMA tile k of this group into stage s: one bulk copy when the batch is dense
        // (stride == N), one per window otherwise (sliding / overlapping windows)
        const uint32_t tile = tile0 + k;
        const uint32_t w0 = tile * WPT;
        uint32_t nw = n_windows - w0;
        if (nw > WPT) nw = WPT;
        const uint32_t bytes = (WPT > 1 && p.stride < N) ? uint32_t((nw - 1) * p.stride + N) * 8u : nw * N * 8u;
        mbar_expect_tx(bar0 + s * 8, bytes);
        const float2 *src;
        if constexpr (MULTI) {
            const uint32_t c = w0 / p.seg_windows;
            src = segs.p[c] + size_t(w0 - c * p.seg_windows) * p.stride;
        } else {
            src = p.iq + size_t(w0) * p.stride;
        }
        if (WPT == 1 || p.stride == N) {
            tma_load_1d(gbase + s * C::STAGE_BYTES, src, bytes, bar0 + s * 8);
        } else if (p.stride < N) {
            // overlapping windows (a sliding search): the tile's windows share samples, ONE copy of the span
            // they cover; window i of the tile then starts stride samples (not N) after window i - 1 in the stage
            tma_load_1d(gbase + s * C::STAGE_BYTES, src, bytes, bar0 + s * 8);
        } else {
            for (uint32_t i = 0; i < nw; ++i)
                tma_load_1d(gbase + s * C::STAGE_BYTES + i * N * 8u, src + size_t(i) * p.stride,
                            N * 8u, bar0 + s * 8);
        }
    };

    if (tg == 0) {
        // the group's own thread 0 initialises its stage barriers and may use them at once;
        // the other threads only touch them after the __syncthreads below
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + s * 8, 1);
        if (tid == 0 && C::CHIRP_SMEM) mbar_init(cbar, 1);
        fence_proxy_async(); // make the inits visible to the async (TMA) proxy
        if (tid == 0 && C::CHIRP_SMEM) {
            // chirp table -> shared memory by TMA (same linear layout as a window)
            mbar_expect_tx(cbar, uint32_t(N) * 8u);
            tma_load_1d(chirp_s, p.chirp, uint32_t(N) * 8u, cbar);
        }
        // With LORAB200_FLAG_INPUT_STABLE the caller guarantees that the IQ batch was not
        // produced by the kernel launched just before this one on the stream (a streaming
        // receiver's double buffering), so the first TMA loads start right here, before the
        // prior grid has drained; the results are still only written after the wait below.
        if (p.early_iq) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                if (uint32_t(s) < my_tiles) issue(s, s);
        }
    }
    // twiddles of the radix-16 passes (registers for the whole persistent loop); their global
    // loads overlap the barrier and the first TMA round trip
    float2 tw1a[3], tw1b[4][3];
    float2 tw2a[3], tw2b[4][3];
    // Thread coordinates of the radix-16 passes.  Two passes: lo1 = t (hi1 = 0).  Three
    // passes: thread t of the window plays (lo, hi) = (t >> 4, t & 15) in pass 1 -- output
    // index lo | y << B0 | hi << (B0+4) -- and (lo, m = t & 15) in pass 2 -- output index
    // tL | y << (B0+4) with tL = lo | m << B0.  The second exchange then only moves data
    // between the 16 threads of one lo-group, which are 16 consecutive lanes of one warp:
    // it needs a __syncwarp, not a group barrier, and each thread writes back exactly the
    // 16 slots it gathered.
    const int lo1 = (C::NP == 3) ? (t >> 4) : t;
    const int hi1 = (C::NP == 3) ? (t & 15) : 0;
    const int tL = (C::NP == 3) ? (lo1 | (hi1 << B0)) : t;   // owner coordinate of the final bins tL + G*y
    load_pass_tw<SF, B0>(p.tw, lo1, tw1a, tw1b);
    if constexpr (C::NP == 3) load_pass_tw<SF, B0 + 4>(p.tw, tL, tw2a, tw2b);

    __syncthreads(); // barrier inits visible CTA-wide (the chirp barrier is shared by all groups)
    // With a stable input batch the only thing that must be ordered after the previous grid is the
    // first global WRITE (the result records, written at a flush): the wait moves down to there and
    // a CTA that became resident early demodulates its windows while the previous grid's stragglers
    // finish.  The debug dumps (written inside the loop) and the walking fine-tune index (p.fine_x16
    // comes from the prefix kernels launched just before) keep the wait here.
    constexpr bool CAN_DEFER = !DBG && FINE != FINE_WALK;
    bool waited = !(CAN_DEFER && p.early_iq);
    if (waited) pdl_wait_prior_grid();
    if (!p.early_iq && tg == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if (uint32_t(s) < my_tiles) issue(s, s);
    }
    if constexpr (C::CHIRP_SMEM) mbar_wait(cbar, 0);

    // per-thread addresses (bytes)
    // pass-0 input: i(e, c) = ilow(e) + (N/R0)*c,  ilow = t (E==1) or ehi*2G + 2t + e0
    const uint32_t wstep = (WPT > 1 && p.stride < N) ? uint32_t(p.stride) : uint32_t(N);   // window pitch inside a stage
    const uint32_t in_off = (uint32_t(wsub) * wstep + uint32_t(E == 1 ? t : 2 * t)) * 8u;
    const uint32_t chirp_off = uint32_t(E == 1 ? t : 2 * t) * 8u;
    const uint32_t wx = xbase + uint32_t(wsub) * C::WSTRIDE * 8u; // this window's exchange area
    // pass-0 scatter: j = x(c) | digit_rev(ilow(e)) << B0.  The e part of ilow lands in bits of j that the
    // t part never touches and that do not carry through the padding (sc0_split_ok checks every t), so
    // sub-FFT e scatters to "sub-FFT 0's address + a compile-time constant": one address register, not E
    static_assert(sc0_split_ok<SF>(), "pass-0 scatter: e offsets are not thread independent");
    const uint32_t sc0 = wx + uint32_t(C::pad(digit_rev4(E == 1 ? t : 2 * t, SF - B0) << B0)) * 8u;
    // pass-1 gather/scatter: j = lo | y << B0 | hi << (B0+4);  pass-2 gather: j = t | y << (B0+4)
    const uint32_t ga1 = wx + uint32_t(C::pad(lo1 | (hi1 << (B0 + 4)))) * 8u;
    const uint32_t ga2 = wx + uint32_t(C::pad(tL)) * 8u;

    int orbit_j = 0, orbit_r = 0;   // FINE_ORBIT: segment / offset of this thread's current window start
    bool orbit_init = false;
    uint32_t raw_cnt = 0;           // raw results parked in rawbuf
    uint32_t raw_w0 = tile0 * WPT;  // window index of rawbuf[0]

    auto flush = [&]() {
        // warp 0 of the group finalises the parked records, one per lane
        group_sync<GROUP>(grp);
        if (!waited) { pdl_wait_prior_grid(); waited = true; }
        if (tg < 32) {
            if (uint32_t(lane) < raw_cnt) {
                const RawResult rr = rawbuf[lane];
                if (p.raw) p.raw[raw_w0 + lane] = rr;
                if (LORAB200_ABL & 8) { if (p.out) { lorab200_result o; o.index = rr.index; o.power = rr.max_mag2; o.power_avg = float(rr.total); o.f_index = rr.left.x + rr.right.y; p.out[raw_w0 + lane] = o; } } // ablation: no dB math
                else
                if (p.out) finalize_device(rr, p.power_scale, &p.out[raw_w0 + lane]);
            }
        }
        group_sync<GROUP>(grp);
        raw_w0 += raw_cnt;
        raw_cnt = 0;
    };

    for (uint32_t k0 = 0; k0 < my_tiles; k0 += STAGES) {
        const uint32_t parity = (k0 / STAGES) & 1u;
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const uint32_t k = k0 + s;
            if (k >= my_tiles) break;
            const uint32_t stage = gbase + uint32_t(s) * C::STAGE_BYTES;
            const uint32_t w = (tile0 + k) * WPT + wsub;     // this thread's window
            const bool live = w < n_windows;

            c64 r[16];
            float2 chg[C::CHIRP_SMEM ? 1 : 16];
            if constexpr (!C::CHIRP_SMEM) {
                // chirp through the read-only path, issued before the wait on the tile
#pragma unroll
                for (int e = 0; e < E; ++e)
#pragma unroll
                    for (int c = 0; c < R0; ++c) {
                        const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
                        chg[e * R0 + c] = __ldg(p.chirp + il + (N / R0) * c);
                    }
            }

            // ---- walking fine-tune index: this thread's 16 table indices ---------------------------
            int fidx[(FINE == FINE_WALK || FINE == FINE_ORBIT) ? 16 : 1];
            float2 fph[FINE == FINE_ORBIT ? 16 : 1];
            if constexpr (FINE == FINE_ORBIT) {
                // (segment, offset) of this thread's window start: located once, then advanced by the WPT
                // windows between consecutive tiles of the group; a window that lies inside one
                // arithmetic progression (all but one in 128/|s|) takes one multiply-add per sample
                if (!orbit_init) { orbit_locate(orb.o, (unsigned long long)w * N, orbit_j, orbit_r); orbit_init = true; }
                else orbit_r += WPT * N;
                (void)orbit_state(orb.o, orbit_j, orbit_r);
                const int seg_y = orb.o.seg[orbit_j].y, seg_c = orb.o.seg[orbit_j].c, os = orb.o.s;
                if (orbit_r + (N - 1) <= seg_c) {
                    // the whole window lies inside one progression (decided once per window, the same for
                    // all its threads): index = base - sample * s, one multiply-add per sample and no
                    // per-sample test
                    const int base = seg_y - orbit_r * os;
#pragma unroll
                    for (int e = 0; e < E; ++e)
#pragma unroll
                        for (int c = 0; c < R0; ++c) {
                            const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
                            fidx[e * R0 + c] = base - (il + (N / R0) * c) * os;
                        }
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e)
#pragma unroll
                        for (int c = 0; c < R0; ++c) {
                            const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
                            int rr = orbit_r + il + (N / R0) * c;
                            if (rr <= seg_c) fidx[e * R0 + c] = seg_y - rr * os;
                            else { int jj = orbit_j; fidx[e * R0 + c] = orbit_state(orb.o, jj, rr); }
                        }
                }
                // the phasors do not depend on the IQ tile: gather them (L2) before waiting for it
#pragma unroll
                for (int i = 0; i < 16; ++i) fph[i] = __ldg(p.fine + fidx[i]);
            }
            if constexpr (FINE == FINE_WALK) {
                group_sync<GROUP>(grp); // the previous tile's gathers from X are done group-wide
                int x = live ? __ldg(p.fine_x16 + size_t(w) * G + t) : 0;
                const uint32_t fx = xbase + uint32_t(wsub) * C::FINE_WSTRIDE * 4u; // this window's index area
                const uint32_t cbase = fx + uint32_t(fine_pad_dyn(16 * t)) * 4u;
                // Almost every chunk lies inside one region of the recurrence, where each step subtracts
                // the same integer (fine_walk.h): then the 16 indices are an arithmetic progression.
                // Chunks that touch a region boundary or the wrap walk the recurrence literally.
                const int xe = x - 15 * p.fine_lin_step;
                const bool linear = p.fine_lin_ok && min(x, xe) >= p.fine_lin_lo && max(x, xe) < p.fine_len;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int v[4];
                    if (linear) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[i] = x - (4 * q + i) * p.fine_lin_step;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { v[i] = x; x = fine_next(x, p.fine_step, p.fine_len); }
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + q * 16u),
                                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
                }
                group_sync<GROUP>(grp);
                // gather in pass-0 order.  j = jc + ilow(t): for every SF the constant part jc and
                // the thread part never carry into each other's 32-block, so the padded address is
                // "thread base + immediate" (tools/layout_model.py checks it and the bank conflicts)
                const uint32_t gb = fx + uint32_t(fine_pad_dyn(E == 1 ? t : 2 * t)) * 4u;
                if constexpr (E == 1) {
#pragma unroll
                    for (int c = 0; c < R0; ++c)
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(fidx[c])
                                     : "r"(gb + uint32_t(fine_pad((N / R0) * c)) * 4u));
                } else {
#pragma unroll
                    for (int eh = 0; eh < E / 2; ++eh)
#pragma unroll
                        for (int c = 0; c < R0; ++c)
                            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                                         : "=r"(fidx[(2 * eh) * R0 + c]), "=r"(fidx[(2 * eh + 1) * R0 + c])
                                         : "r"(gb + uint32_t(fine_pad(eh * 2 * G + (N / R0) * c)) * 4u));
                }
            }

            mbar_wait(bar0 + s * 8, parity);

            // ---- pass 0: load (linear layout) -------------------------------------------------------
            {
                const uint32_t in0 = stage + in_off;
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
            }
            // the stage is now in registers: hand it back to the TMA for tile k + STAGES.  Normally
            // at once (one group barrier, then thread 0 issues the refill).  SF10 and SF12 instead
            // issue the refill after the barrier that follows the exchange-0 scatter (every thread
            // has read the stage by then), which saves one named barrier per tile: measured +2 % /
            // +3 %; SF11 (one stage, two groups per CTA) loses 7 % to the later prefetch and keeps
            // the early hand-back.  FINE_WALK keeps the early barrier: it also separates the index
            // gather from the scatter.
            // (The fence.proxy.async before the refill stays: a build without it failed a parity
            // test on the GPU and was not faster.)
            constexpr bool LATE_REFILL = (SF == 10 || SF == 12) && FINE != FINE_WALK;
            if constexpr (!LATE_REFILL) {
                group_sync<GROUP>(grp);
                if (tg == 0 && k + STAGES < my_tiles) { fence_proxy_async(); issue(k + STAGES, s); }
            }

            // ---- dechirp ------------------------------------------------------------------------
            if constexpr (C::CHIRP_SMEM) {
                const uint32_t c0 = chirp_s + chirp_off;
                if constexpr (E == 1) {
#pragma unroll
                    for (int c = 0; c < R0; ++c)
                        r[c] = pcx_mul(r[c], cunpack(lds64(c0 + uint32_t((N / R0) * c) * 8u)));
                } else {
#pragma unroll
                    for (int eh = 0; eh < E / 2; ++eh)
#pragma unroll
                        for (int c = 0; c < R0; ++c) {
                            c64 ca, cb;
                            lds128(c0 + uint32_t(eh * 2 * G + (N / R0) * c) * 8u, ca, cb);
                            r[(2 * eh) * R0 + c] = pcx_mul(r[(2 * eh) * R0 + c], cunpack(ca));
                            r[(2 * eh + 1) * R0 + c] = pcx_mul(r[(2 * eh + 1) * R0 + c], cunpack(cb));
                        }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = pcx_mul(r[i], chg[i]);
            }
            if constexpr (FINE0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = pcx_mul(r[i], fine_phasor);
            }
            if constexpr (FINE == FINE_WALK) {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = pcx_mul(r[i], __ldg(p.fine + fidx[i]));
            }
            if constexpr (FINE == FINE_ORBIT) {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = pcx_mul(r[i], fph[i]);
            }
            if constexpr (DBG) {
                if (p.dec && live) {
#pragma unroll
                    for (int e = 0; e < E; ++e)
#pragma unroll
                        for (int c = 0; c < R0; ++c) {
                            const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
                            p.dec[size_t(w) * N + il + (N / R0) * c] = cunpack(r[e * R0 + c]);
                        }
                }
            }

            // ---- pass 0 butterflies (deepest kissfft stages; tw[0] multiplies skipped) -------------
#pragma unroll
            for (int e = 0; e < E; ++e) {
                c64 *v = r + e * R0;
                if constexpr (B0 == 1) {
                    p_bfly2(v[0], v[1]);
                } else if constexpr (B0 == 2) {
                    p_bfly4(v[0], v[1], v[2], v[3]);
                } else if constexpr (B0 == 3) {
                    // radix-2 (m=1) on pairs (c, c+4), then radix-4 (m=2) over c = q + 4k
#pragma unroll
                    for (int a = 0; a < 4; ++a) p_bfly2(v[a], v[a + 4]);
                    p_bfly4(v[0], v[1], v[2], v[3]);                       // k = 0: tw[0]
                    v[5] = pcx_mul(v[5], p0tw.w[0][0]);
                    v[6] = pcx_mul(v[6], p0tw.w[1][0]);
                    v[7] = pcx_mul(v[7], p0tw.w[2][0]);
                    p_bfly4(v[4], v[5], v[6], v[7]);
                } else {
                    // radix-4 (m=1) over qa for each qb (slots qb + 4*qa), then radix-4 (m=4)
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb) p_bfly4(v[qb], v[qb + 4], v[qb + 8], v[qb + 12]);
                    p_bfly4(v[0], v[1], v[2], v[3]);                       // k = 0
#pragma unroll
                    for (int kk = 1; kk < 4; ++kk) {
                        v[4 * kk + 1] = pcx_mul(v[4 * kk + 1], p0tw.w[0][kk - 1]);
                        v[4 * kk + 2] = pcx_mul(v[4 * kk + 2], p0tw.w[1][kk - 1]);
                        v[4 * kk + 3] = pcx_mul(v[4 * kk + 3], p0tw.w[2][kk - 1]);
                        p_bfly4(v[4 * kk], v[4 * kk + 1], v[4 * kk + 2], v[4 * kk + 3]);
                    }
                }
            }

            // ---- exchange 0.  The previous tile's reads of X are over: warp-sized groups passed the
            // __syncwarp above, larger groups the cross-warp reduction barrier of the previous tile -----
#pragma unroll
            for (int e = 0; e < E; ++e)
#pragma unroll
                for (int c = 0; c < R0; ++c)
                    sts64(sc0 + uint32_t(sc0_eoff<SF>(e) + C::pad(slot_to_x<B0>(c))) * 8u, r[e * R0 + c]);
            group_sync<GROUP>(grp);
            if constexpr (LATE_REFILL) {
                if (tg == 0 && k + STAGES < my_tiles) { fence_proxy_async(); issue(k + STAGES, s); }
            }
#pragma unroll
            for (int y = 0; y < 16; ++y) r[y] = lds64(ga1 + uint32_t(C::pad(y << B0)) * 8u);

            // ---- pass 1 (and exchange 1 + pass 2) -----------------------------------------------------
            radix16(r, tw1a, tw1b);
            if constexpr (C::NP == 3) {
                // exchange 1, in place: every thread overwrites the 16 slots it has just gathered
                // (no other thread reads them in exchange 0), and the gather below only touches
                // slots written by the 16 lanes of this thread's lo-group
#pragma unroll
                for (int y = 0; y < 16; ++y) sts64(ga1 + uint32_t(C::pad(y << B0)) * 8u, r[y]);
                __syncwarp();
#pragma unroll
                for (int y = 0; y < 16; ++y) r[y] = lds64(ga2 + uint32_t(C::pad(y << (B0 + 4))) * 8u);
                radix16(r, tw2a, tw2b);
            }

            // ---- peak scan: thread holds bins j = t + G*y, y ascending (LoRaDetector.hpp:36-48) ----------
            // Tournament over the 16 bins instead of a 16-long compare chain: at every node the
            // lower bin wins unless the higher one is strictly larger, which is exactly what the
            // sequential strict-'>' scan yields; a final strict compare against 0 reproduces the
            // "running maximum starts at 0" (an all-zero or all-NaN window keeps index 0).
            float m2[16];
#pragma unroll
            for (int y = 0; y < 16; ++y) m2[y] = pcx_mag2(r[y]);

            float bv; int by;
            {
                float v8[8]; int i8[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const bool hi = m2[2 * q + 1] > m2[2 * q];
                    v8[q] = hi ? m2[2 * q + 1] : m2[2 * q];
                    i8[q] = hi ? 2 * q + 1 : 2 * q;
                }
                float v4[4]; int i4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool hi = v8[2 * q + 1] > v8[2 * q];
                    v4[q] = hi ? v8[2 * q + 1] : v8[2 * q];
                    i4[q] = hi ? i8[2 * q + 1] : i8[2 * q];
                }
                const bool h0 = v4[1] > v4[0], h1 = v4[3] > v4[2];
                const float va = h0 ? v4[1] : v4[0], vb = h1 ? v4[3] : v4[2];
                const int ia = h0 ? i4[1] : i4[0], ib = h1 ? i4[3] : i4[2];
                const bool hh = vb > va;
                bv = hh ? vb : va;
                by = hh ? ib : ia;
                if (!(bv > 0.0f)) { bv = 0.0f; by = 0; }
            }
            // double sum of the float |bin|^2 terms, four partial chains
            double s0 = double(m2[0]), s1 = double(m2[1]), s2 = double(m2[2]), s3 = double(m2[3]);
#pragma unroll
            for (int y = 4; y < 16; y += 4) {
                s0 += double(m2[y]); s1 += double(m2[y + 1]); s2 += double(m2[y + 2]); s3 += double(m2[y + 3]);
            }
            double sum = (s0 + s1) + (s2 + s3);
            if (LORAB200_ABL & 1) sum = double(m2[0] + m2[5]);   // ablation: no double sum
            if (LORAB200_ABL & 2) { bv = fmaxf(m2[3], m2[7]); by = int(__float_as_uint(m2[1]) & 15u); } // ablation: no tournament
            int bi = tL + G * by;
            if constexpr (DBG) {
                if (p.fft && live) {
#pragma unroll
                    for (int y = 0; y < 16; ++y) p.fft[size_t(w) * N + tL + G * y] = cunpack(r[y]);
                }
            }
            // reduce over the window's G threads
            constexpr int WIDTH = G < 32 ? G : 32;
            if constexpr (G >= 32) {
                // the window spans whole warps: REDUX.  bv is >= +0 and never NaN here, so the
                // unsigned order of its bit pattern is the float order; among equal maxima the
                // smallest index wins (the sequential scan's first-max rule)
                const uint32_t bits = __float_as_uint(bv);
                const uint32_t mx = __reduce_max_sync(0xffffffffu, bits);
                bi = int(__reduce_min_sync(0xffffffffu, bits == mx ? uint32_t(bi) : 0xffffffffu));
                bv = __uint_as_float(mx);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            } else {
#pragma unroll
                for (int o = WIDTH / 2; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    argmax_merge(bv, bi, ov, oi);
                }
            }
            if constexpr (G > 32) {
                // cross-warp: warp leaders publish one 16-byte partial {sum, value bits, index};
                // lane l of every warp picks up partial l mod K and the K partials are folded with
                // the same REDUX / xor-shuffle pattern (every aligned block of K lanes holds all K)
                constexpr int K = G / 32;
                const int wig = tg >> 5;
                const uint32_t rbase = smem_u32(red_sum);
                if (lane == 0)
                    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(rbase + uint32_t(wig) * 16u),
                                 "l"(__double_as_longlong(sum)),
                                 "l"((unsigned long long)__float_as_uint(bv) | ((unsigned long long)uint32_t(bi) << 32))
                                 : "memory");
                group_sync<GROUP>(grp);
                long long ps; unsigned long long pk;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(ps), "=l"(pk)
                             : "r"(rbase + uint32_t(lane & (K - 1)) * 16u));
                sum = __longlong_as_double(ps);
                const uint32_t bits = uint32_t(pk), idx = uint32_t(pk >> 32);
                const uint32_t mx = __reduce_max_sync(0xffffffffu, bits);
                bi = int(__reduce_min_sync(0xffffffffu, bits == mx ? idx : 0xffffffffu));
                bv = __uint_as_float(mx);
#pragma unroll
                for (int o = K / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            }
            // neighbours of the peak, (bi-1) mod N and (bi+1) mod N, live in other threads' registers
            // (cyclic neighbours, LoRaDetector.hpp:56-57: N is a power of two, so the wrap is a mask)
            const unsigned jl = (unsigned(bi) + unsigned(N - 1)) & unsigned(N - 1);
            const unsigned jr = (unsigned(bi) + 1u) & unsigned(N - 1);
            const int tl = int(jl & unsigned(G - 1)), yl = int(jl >> (SF - 4));
            const int tr = int(jr & unsigned(G - 1)), yr = int(jr >> (SF - 4));
            // (only the warps that hold one of the two owner lanes run the select tree)
            c64 nbv = r[0];
            if (LORAB200_ABL & 4) nbv = r[5];                    // ablation: no neighbour select
            else
            if (G <= 32 || __any_sync(0xffffffffu, tL == tl || tL == tr)) nbv = pick16(r, tL == tl ? yl : yr);
            // the record of this window is assembled in shared memory by the three lanes that
            // own its pieces; it is read back at the next flush (after a group-wide sync)
            if (live) {
                RawResult *slot = rawbuf + raw_cnt + wsub;
                if (tL == tl) slot->left = cunpack(nbv);
                if (tL == tr) slot->right = cunpack(nbv);
                if (t == 0) { slot->index = bi; slot->max_mag2 = bv; slot->total = sum; }
            }
            {
                uint32_t nw = n_windows - (tile0 + k) * WPT;
                raw_cnt += nw > WPT ? WPT : nw;
            }
            if (raw_cnt + WPT > C::RAWCAP) flush();
        }
    }
    if (raw_cnt) flush();
}

// ---- host side -------------------------------------------------------------------------
#ifndef LORAB200_KERNEL_ONLY // (tools: compile a single instantiation for SASS inspection)
inline bool fused_applicable(int sf, const DemodParams &p)
{
    if (sf < 7 || sf > 12) return false;
    const unsigned long long n = 1ull << sf;
    // (multi-segment batches: the caller has checked every segment pointer, p.iq is the first)
    if ((reinterpret_cast<uintptr_t>(p.iq) & 15u) != 0) return false; // TMA needs 16-byte aligned sources:
    if ((p.stride & 1ull) != 0 && p.n_windows > 1) return false;      // base aligned and an even sample stride
    if (p.n_windows >= (1ull << 31)) return false;
    if (p.seg_windows != 0) {
        // multi-segment: tiles must not straddle segments; no debug dumps, no fine-tune
        const unsigned wpt = n >= 512 ? 1u : unsigned(512 / n);
        if (p.seg_windows % wpt != 0) return false;
        if (p.fft != nullptr || p.dec != nullptr || p.fine != nullptr) return false;
    }
    // walking fine index together with the fft/dec debug dumps: simple kernel
    if (p.fine != nullptr && p.fine_step != 0.0f && (p.fft != nullptr || p.dec != nullptr)) return false;
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

template <int SF, bool DBG, int FINE, bool MULTI = false>
cudaError_t launch_fused_v(const DemodParams &p, const Pass0Tw &tw0, float2 phasor, cudaStream_t s, int sm_count,
                           const SegList<MULTI> &segs = SegList<MULTI>(), const OrbitArg<FINE> &orb = OrbitArg<FINE>())
{
    using C = FusedCfg<SF>;
#ifdef LORAB200_EXP_1CTA   // occupancy experiment: one CTA per SM (tools/build_variant.sh)
    constexpr int kSmem = 120 * 1024, kCtasPerSm = 1;
#else
    constexpr int kSmem = C::SMEM_BYTES, kCtasPerSm = C::CTAS_PER_SM;
#endif
    static SmemAttrOnce smem_attr; // per (instantiation, device)
    if (cudaError_t e = smem_attr.ensure(demod_fused_kernel<SF, DBG, FINE, MULTI>, kSmem); e != cudaSuccess) return e;
    const unsigned long long n_tiles = (p.n_windows + C::WPT - 1) / C::WPT;
    unsigned long long ctas = (n_tiles + C::GROUPS - 1) / C::GROUPS;
    const unsigned long long cap = (unsigned long long)kCtasPerSm * (unsigned long long)sm_count; // persistent
    if (ctas > cap) ctas = cap;
    // Programmatic dependent launch: consecutive demod launches on a stream may overlap
    // this kernel's prologue (barrier init, table loads) with the previous kernel's tail;
    // the kernel itself waits (griddepcontrol.wait) before touching the batch or the results.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(ctas));
    cfg.blockDim = dim3(C::CTA);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, demod_fused_kernel<SF, DBG, FINE, MULTI>, p, tw0, phasor, segs, orb);
}

template <int SF>
cudaError_t launch_fused_sf(const DemodParams &p, const Pass0Tw &tw0, int fine, float2 phasor,
                            cudaStream_t s, int sm_count, const FineOrbit *orbit)
{
    const bool dbg = p.fft != nullptr || p.dec != nullptr;
    if (fine == FINE_ORBIT) {
        if (dbg || !orbit) return cudaErrorInvalidValue;
        OrbitArg<FINE_ORBIT> oa;
        oa.o = *orbit;
        return launch_fused_v<SF, false, FINE_ORBIT>(p, tw0, phasor, s, sm_count, SegList<false>(), oa);
    }
    if (fine == FINE_WALK) return dbg ? cudaErrorInvalidValue : launch_fused_v<SF, false, FINE_WALK>(p, tw0, phasor, s, sm_count);
    if (dbg) return fine == FINE_CONST ? launch_fused_v<SF, true, FINE_CONST>(p, tw0, phasor, s, sm_count)
                                       : launch_fused_v<SF, true, FINE_NONE>(p, tw0, phasor, s, sm_count);
    return fine == FINE_CONST ? launch_fused_v<SF, false, FINE_CONST>(p, tw0, phasor, s, sm_count)
                              : launch_fused_v<SF, false, FINE_NONE>(p, tw0, phasor, s, sm_count);
}

// multi-segment batch (detector-only form)
inline cudaError_t launch_fused_multi(int sf, const DemodParams &p, const Pass0Tw &tw0, const SegList<true> &segs,
                                      cudaStream_t s, int sm_count)
{
    const float2 one = make_float2(1.f, 0.f);
    switch (sf) {
    case 7: return launch_fused_v<7, false, FINE_NONE, true>(p, tw0, one, s, sm_count, segs);
    case 8: return launch_fused_v<8, false, FINE_NONE, true>(p, tw0, one, s, sm_count, segs);
    case 9: return launch_fused_v<9, false, FINE_NONE, true>(p, tw0, one, s, sm_count, segs);
    case 10: return launch_fused_v<10, false, FINE_NONE, true>(p, tw0, one, s, sm_count, segs);
    case 11: return launch_fused_v<11, false, FINE_NONE, true>(p, tw0, one, s, sm_count, segs);
    case 12: return launch_fused_v<12, false, FINE_NONE, true>(p, tw0, one, s, sm_count, segs);
    default: return cudaErrorInvalidValue;
    }
}

// fine: FINE_NONE, FINE_CONST (multiply by `phasor` = fine[idx0], step == 0), FINE_ORBIT (`orbit`
// describes the walk) or FINE_WALK (p.fine_x16 holds the index at every 16th sample)
inline cudaError_t launch_fused(int sf, const DemodParams &p, const Pass0Tw &tw0, int fine, float2 phasor,
                                cudaStream_t s, int sm_count, const FineOrbit *orbit = nullptr)
{
    switch (sf) {
    case 7: return launch_fused_sf<7>(p, tw0, fine, phasor, s, sm_count, orbit);
    case 8: return launch_fused_sf<8>(p, tw0, fine, phasor, s, sm_count, orbit);
    case 9: return launch_fused_sf<9>(p, tw0, fine, phasor, s, sm_count, orbit);
    case 10: return launch_fused_sf<10>(p, tw0, fine, phasor, s, sm_count, orbit);
    case 11: return launch_fused_sf<11>(p, tw0, fine, phasor, s, sm_count, orbit);
    case 12: return launch_fused_sf<12>(p, tw0, fine, phasor, s, sm_count, orbit);
    default: return cudaErrorInvalidValue;
    }
}
#endif // LORAB200_KERNEL_ONLY

} // namespace lorab200
