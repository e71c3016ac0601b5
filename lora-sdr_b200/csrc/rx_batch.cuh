// rx_batch.cuh -- the receiver kernel, second generation (row f3): sync acquisition and symbol
// extraction on the device with the fused kernel's register FFT and SPECULATIVE batches.
//
// One CTA per channel runs the reference demod's work() sequence (LoRaDemod.cpp:145-327) over its
// stream, like rx_kernel.cuh, but a round demodulates up to B windows at once -- the windows at
// off, off+N, off+2N, ... under the state the channel is in now (chirp table, fine-tune step,
// fine-tune index continuing from window to window) -- with the register-resident kissfft passes of
// demod_fused.cuh (16 points per thread, one group of max(32, N/16) threads per window, padded
// exchange buffers), B = 32 / 16 / 8 / 4 / 2 / 1 windows for SF 7..12 in a 256-thread CTA.  Thread 0
// then walks the results through the state machine of rx_fsm.h for as long as the speculation holds:
// a step must consume exactly N samples and leave (table, _finefreqError, _fineTuneIndex) as the next
// window assumed.  That is every window of STATE_DATASYMBOLS (positions known, LoRaDemod.cpp:289),
// every squelched window of an idle STATE_FRAMESYNC, and the look-ahead window of the sync word
// (:189-206 demodulates it with the same table and a continuing index); a preamble window that
// re-times the stream (`total = N - value`, :219) ends the walk and the next round starts at the new
// offset.  In the phases that are inherently one window at a time the round only demodulates two.
//
// Numerics are the fused kernel's: the reference's float operand graph on the packed FP32 pipe;
// records finalised with finalize_device.  Outputs are identical to rx_kernel.cuh's (the simple
// routine), which stays as the cross-check (LORAB200_FLAG_SIMPLE_ONLY); tests/test_gpu_rx.py
// compares both with the unmodified reference block.
#pragma once
#include "demod_fused.cuh"
#include "rx_kernel.cuh"

namespace lorab200 {

// -DLORAB200_RX_PROF=1: cycle accounting of the one-warp (solo) rounds; =2: of every round of at most two windows
// (the sequential phases at any SF), as thread 0 sees them.  Printed by the kernel: a tuning aid, not part of
// the product build.  (prof_on is latched when a round starts, so that the time of rounds that are not accounted
// is not attributed to the next marker.)
#ifdef LORAB200_RX_PROF
#define RXPROF(i) do { if (prof_on && tid == 0) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); \
                       prof_acc[i] += t_ - prof_last; prof_last = t_; } } while (0)
#else
#define RXPROF(i) do { } while (0)
#endif

// Cold paths kept out of line: the receiver kernel's hot loop should fit the instruction caches (a build
// with everything inlined ran many-channel launches at half the speed).
__device__ __noinline__ int fine_advance_cold(const FineWalkDev *fw, int x, unsigned long long n)
{
    return fine_advance_dev(*fw, x, n);
}
__device__ __noinline__ void fine_walk_make_cold(FineWalkDev *fw, FineWalk32 *fw32, int len, float d)
{
    *fw = fine_walk_make_dev(len, d);
    *fw32 = fine_walk32_make(*fw);
}

template <int SF>
struct RxBatchCfg {
    using C = FusedCfg<SF>;
    static constexpr int N = C::N, G = C::G, GROUP = C::GROUP, WPT = C::WPT;
    static constexpr int CTA = 256;
    static constexpr int GROUPS = CTA / GROUP;
    static constexpr int B = GROUPS * WPT;                 // windows per round
    static constexpr int BR = B < 2 ? 2 : B;               // result slots (SF12: the window and its look-ahead)
    static constexpr int X_BYTES = C::X_BYTES;
    static constexpr int RED_BYTES = 256;
    static constexpr int GROUP_BYTES = X_BYTES + RED_BYTES;
    static constexpr int SMEM_BYTES = GROUPS * GROUP_BYTES;
};

struct RxRoundRes {
    lorab200_result r;
    int idx_after;      // _fineTuneIndex after the window's N samples (== what the next window assumed)
    float max_mag2;     // raw peak |bin|^2 of the window (for a host-side, libm-exact "power" signal)
    double total;       // raw sum of |bin|^2
};

template <int SF>
__global__ void __launch_bounds__(256, 2) rx_batch_kernel(const RxParams p, const Pass0Tw p0tw)
{
    using K = RxBatchCfg<SF>;
    using C = FusedCfg<SF>;
    constexpr int N = C::N, B0 = C::B0, R0 = C::R0, E = C::E, G = C::G, GROUP = C::GROUP, WPT = C::WPT;
    constexpr int B = K::B;

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ RxState st;
    __shared__ FineWalkDev fw;
    __shared__ FineWalk32 fw32;
    __shared__ int s_spec_ok, s_spec_m;     // one-warp rounds: closed-form guess of the chunk starts allowed / its modulus
    __shared__ int s_after[K::BR];           // _fineTuneIndex after each window of the round (its last chunk's walk)
    __shared__ RawResult rawres[K::BR];
    __shared__ RxRoundRes res[K::BR];
    __shared__ unsigned long long off;
    __shared__ int s_go, s_nwin, s_table, s_x0, s_walk, npk, s_lin_ok, s_lin_step, s_lin_lo, s_force2, s_solo, s_need_f;
    __shared__ float s_d, s_fw_d;
    __shared__ float2 s_phasor;
#ifdef LORAB200_RX_PROF
    __shared__ long long prof_acc[12], prof_last, prof_rounds;
    __shared__ int prof_on;
    if (threadIdx.x == 0) { for (int i = 0; i < 12; ++i) prof_acc[i] = 0; prof_rounds = 0; prof_on = 0; asm volatile("mov.u64 %0, %%clock64;" : "=l"(prof_last) :: "memory"); }
#endif

    const int tid = threadIdx.x;
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int grp = GROUP <= 32 ? warp_u : warp_u / (GROUP / 32);
    const int tg = tid % GROUP;
    const int wsub = tg / G;
    const int t = tg % G;
    const int lane = tid & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t xbase = smem_base + uint32_t(grp) * K::GROUP_BYTES;
    double *red_sum = reinterpret_cast<double *>(smem + grp * K::GROUP_BYTES + K::X_BYTES);

    // twiddles of the radix-16 passes, in registers for the whole kernel (as in demod_fused.cuh)
    float2 tw1a[3], tw1b[4][3];
    float2 tw2a[3], tw2b[4][3];
    const int lo1 = (C::NP == 3) ? (t >> 4) : t;
    const int hi1 = (C::NP == 3) ? (t & 15) : 0;
    const int tL = (C::NP == 3) ? (lo1 | (hi1 << B0)) : t;
    load_pass_tw<SF, B0>(p.tw, lo1, tw1a, tw1b);
    if constexpr (C::NP == 3) load_pass_tw<SF, B0 + 4>(p.tw, tL, tw2a, tw2b);

    const uint32_t wx = xbase + uint32_t(wsub) * C::WSTRIDE * 8u;
    uint32_t sc0[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
        sc0[e] = wx + uint32_t(C::pad(digit_rev4(il, SF - B0) << B0)) * 8u;
    }
    const uint32_t ga1 = wx + uint32_t(C::pad(lo1 | (hi1 << (B0 + 4)))) * 8u;
    const uint32_t ga2 = wx + uint32_t(C::pad(tL)) * 8u;

    for (unsigned ch = blockIdx.x; ch < p.n_channels; ch += gridDim.x) {
        const float2 *in = p.iq + (unsigned long long)ch * p.chan_stride;
        int16_t *slots = p.symbols + (size_t(ch) * p.max_packets) * p.cfg.mtu;
        RxPacketInfo *info = p.info + size_t(ch) * p.max_packets;
        if (tid == 0) { st = p.state[ch]; off = 0; npk = 0; s_fw_d = 0.0f; s_force2 = 0; }
        __syncthreads();
        if (st.phase == RX_DATASYMBOLS && p.max_packets > 0)
            for (int k = tid; k < st.sym_count; k += K::CTA) slots[k] = p.partial[size_t(ch) * p.cfg.mtu + k];
        __syncthreads();

        // g^n(x) of the fine-tune recurrence: 32-bit region walk when the step allows it
        auto adv = [&](int x, unsigned long long n) -> int {
            return (fw32.ok && n < (1ull << 28)) ? fine_advance32(fw32, x, unsigned(n)) : fine_advance_cold(&fw, x, n);
        };
        // thread 0 plans a round: how many windows, under which state (shared s_* variables)
        auto plan = [&]() {
            const unsigned long long left = p.n_samples - off;
            s_go = (left >= 2ull * N) && (unsigned(npk) < p.max_packets);
            if (s_go) {
                unsigned long long n_avail = left / N;                    // whole windows from `off`
                // SF12 holds ONE window per pass: the look-ahead window of a sync word is demodulated
                // only when the walk asked for it (s_force2), as a second pass
                const int cap = (B == 1) ? (s_force2 ? 2 : 1) : B;
                int nwin = int(n_avail > (unsigned long long)cap ? cap : n_avail);
                // phases in which a result moves the next window: the window itself + its look-ahead
                const bool idle = st.phase == RX_FRAMESYNC && st.ffe == 0.0f && st.fine_index == 0;
                const bool sequential = st.phase != RX_DATASYMBOLS && !idle;
                if (sequential && nwin > 2) nwin = 2;
                s_nwin = nwin;
                // one-window-at-a-time phases of a window that fits one warp: warp 0 runs them alone, without
                // CTA barriers (the other warps wait at the top of the loop)
                s_solo = (GROUP == 32 && WPT >= 2 && sequential) ? 1 : 0;
                s_need_f = (st.phase == RX_FRAMESYNC) ? 1 : 0;   // fIndex only feeds _finefreqError in FRAMESYNC (:220)
                s_table = st.table;
                s_x0 = st.fine_index;
                const float d = __fmul_rn(st.ffe, 128.0f);                // _finefreqError * _fineSteps
                s_d = d;
                s_walk = (d != 0.0f);
                if (s_walk && d != s_fw_d) {
                    fine_walk_make_cold(&fw, &fw32, p.fine_len, d);
                    s_fw_d = d;
                    const bool kp_none = fw.Kp == 0x7fffffffffffffffll;
                    s_lin_ok = (fw.integer || kp_none) ? 1 : 0;
                    s_lin_step = int(fw.integer ? fw.D : fw.D + 1);
                    s_lin_lo = fw.integer ? 0 : int(fw.D + 1 > 0 ? fw.D + 1 : 0);
                    s_spec_ok = (s_lin_ok && s_lin_step != 0 && s_lin_step > -(1 << 17) && s_lin_step < (1 << 17)) ? 1 : 0;
                    s_spec_m = (!fw.integer && s_lin_step > 0) ? p.fine_len + 1 : p.fine_len;
                }
                if (!s_walk) s_phasor = p.fine[s_x0];
            }
        };
        auto compute = [&]() {
            const int nwin = s_nwin;
            const bool walking = s_walk != 0;
            const float2 *chirp = s_table == LORAB200_TABLE_DOWN ? p.down : p.up;
            const int passes = (B == 1) ? nwin : 1;

            // (groups whose windows are all beyond nwin sit the round out: in the one-window-at-a-time
            //  phases that is every group but the first, which then has the SM's issue slots to itself)
            const bool group_live = (B == 1) || grp * WPT < nwin;
            for (int pass = 0; group_live && pass < passes; ++pass) {
                const int k = (B == 1) ? pass : grp * WPT + wsub;             // this thread's window in the round
                const bool live = k < nwin;
                const float2 *win = in + off + (unsigned long long)k * N;

                RXPROF(0);   // (top of a pass: plan -> here)
                // ---- fine-tune indices of this thread's 16 pass-0 samples -----------------------------
                int fidx[16];
                if (walking) {
                    const uint32_t fx = xbase + uint32_t(wsub) * C::FINE_WSTRIDE * 4u;
                    const uint32_t cbase = fx + uint32_t(fine_pad_dyn(16 * t)) * 4u;
                    // walk this thread's 16-sample chunk from state x, park the 16 indices, return the state
                    // after the chunk.  Inside the constant-step region of the recurrence (fine_walk.h) a step
                    // is "minus step, then the reference's wrap"; everything else is the literal float step.
                    auto chunk = [&](int x) -> int {
                        const int xe = x - 15 * s_lin_step;
                        const bool linear = s_lin_ok && min(x, xe) >= s_lin_lo && max(x, xe) < p.fine_len;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            int v[4];
                            if (linear) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) v[i] = x - (4 * q + i) * s_lin_step;
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    v[i] = x;
                                    if (s_lin_ok && x >= s_lin_lo) {
                                        x -= s_lin_step;
                                        if (x < 0) x += p.fine_len; else if (x >= p.fine_len) x -= p.fine_len;
                                    } else {
                                        x = fine_next(x, s_d, p.fine_len);
                                    }
                                }
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + q * 16u),
                                         "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
                        }
                        if (linear) {
                            // all 16 states of the chunk are inside the region: the 16th step is "minus step" too
                            x = x - 16 * s_lin_step;
                            if (x < 0) x += p.fine_len; else if (x >= p.fine_len) x -= p.fine_len;
                        }
                        return x;
                    };
                    const unsigned long long n0 = (unsigned long long)k * N + 16ull * t;   // samples before the chunk
                    // One-warp rounds: every chunk of the round is walked by a lane of this warp, so a GUESS of
                    // the chunk starts can be checked exactly -- lane 0 starts from the true state and every
                    // chunk must end where the next lane started.  The guess: while the walk only meets the
                    // constant-step region and the plain wrap, the state after n samples is
                    // (x0 - n * step) mod M, M = len + 1 for a fractional positive step (the one literal step
                    // per wrap of the table subtracts floor(d), not floor(d) + 1), M = len otherwise.  A failed
                    // check (a state in the rounding band, index 0 / D, ...) takes the region walk (attempt 1).
                    int xend = 0;
                    const bool guess = (GROUP == 32) && s_solo && s_spec_ok;
#pragma unroll 1
                    for (int attempt = guess ? 0 : 1; attempt < 2; ++attempt) {
                        int x;
                        if (attempt == 0) {
                            int m = (s_x0 - int(n0) * s_lin_step) % s_spec_m;
                            if (m < 0) m += s_spec_m;
                            x = live ? m : 0;
                        } else {
                            x = live ? adv(s_x0, n0) : 0;
                        }
                        xend = chunk(x);
                        if (attempt == 0) {
                            const int nx = __shfl_down_sync(0xffffffffu, x, 1);
                            const bool has_succ = live && (t < G - 1 || k + 1 < nwin);
                            if (__all_sync(0xffffffffu, !has_succ || xend == nx)) break;
                        }
                    }
                    // the thread that walked the window's last chunk holds the index after the window (what the
                    // state machine carries on, LoRaDemod.cpp:160-162 after N samples)
                    if (live && t == G - 1) s_after[k] = xend;
                    group_sync<GROUP>(grp);
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
                    group_sync<GROUP>(grp);   // the index area is the exchange buffer: gathers done before the scatter
                }

                RXPROF(1);   // walking index
                // the next step starts somewhere in (off, off + 2N]: whatever it is, it needs samples up to 3N
                // ahead; pull the part this round does not read into L1 while it computes
                // (the one-warp rounds of SF7 / SF8 and the one-window passes of SF12, where it measured +7 %; at SF9 /
                //  SF10 the same prefetch in the CTA-wide sequential rounds measured -4 ... -6 % and is left out)
                if ((s_solo || (B == 1 && nwin <= 2)) && k == 0 && t * 16 < N && off + 3ull * N + 16 <= p.n_samples)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(win + 2 * N + t * 16));
                // ---- load (global, 8-byte aligned), dechirp, fine-tune multiply (LoRaDemod.cpp:157-159) ----
                c64 r[16];
#pragma unroll
                for (int e = 0; e < E; ++e)
#pragma unroll
                    for (int c = 0; c < R0; ++c) {
                        const int il = (E == 1) ? t : ((e >> 1) * 2 * G + 2 * t + (e & 1));
                        const int i = il + (N / R0) * c;
                        const float2 v = live ? __ldg(win + i) : make_float2(0.f, 0.f);
                        c64 d = pcx_mul(cpack(v.x, v.y), __ldg(chirp + i));
                        d = pcx_mul(d, walking ? __ldg(p.fine + fidx[e * R0 + c]) : s_phasor);
                        r[e * R0 + c] = d;
                    }

#ifdef LORAB200_RX_PROF
                { const float2 f0 = cunpack(r[0]), f1 = cunpack(r[15]); if (f0.x == 1.2345e30f && f1.y == 3.3e30f) s_go = 0; }  // wait for the loads
#endif
                RXPROF(2);   // loads + dechirp
                // ---- pass 0 butterflies (demod_fused.cuh) ---------------------------------------------
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    c64 *v = r + e * R0;
                    if constexpr (B0 == 1) {
                        p_bfly2(v[0], v[1]);
                    } else if constexpr (B0 == 2) {
                        p_bfly4(v[0], v[1], v[2], v[3]);
                    } else if constexpr (B0 == 3) {
#pragma unroll
                        for (int a = 0; a < 4; ++a) p_bfly2(v[a], v[a + 4]);
                        p_bfly4(v[0], v[1], v[2], v[3]);
                        v[5] = pcx_mul(v[5], p0tw.w[0][0]);
                        v[6] = pcx_mul(v[6], p0tw.w[1][0]);
                        v[7] = pcx_mul(v[7], p0tw.w[2][0]);
                        p_bfly4(v[4], v[5], v[6], v[7]);
                    } else {
#pragma unroll
                        for (int qb = 0; qb < 4; ++qb) p_bfly4(v[qb], v[qb + 4], v[qb + 8], v[qb + 12]);
                        p_bfly4(v[0], v[1], v[2], v[3]);
#pragma unroll
                        for (int kk = 1; kk < 4; ++kk) {
                            v[4 * kk + 1] = pcx_mul(v[4 * kk + 1], p0tw.w[0][kk - 1]);
                            v[4 * kk + 2] = pcx_mul(v[4 * kk + 2], p0tw.w[1][kk - 1]);
                            v[4 * kk + 3] = pcx_mul(v[4 * kk + 3], p0tw.w[2][kk - 1]);
                            p_bfly4(v[4 * kk], v[4 * kk + 1], v[4 * kk + 2], v[4 * kk + 3]);
                        }
                    }
                }
                // ---- exchange 0, pass 1 (exchange 1, pass 2) -----------------------------------------
#pragma unroll
                for (int e = 0; e < E; ++e)
#pragma unroll
                    for (int c = 0; c < R0; ++c)
                        sts64(sc0[e] + uint32_t(C::pad(slot_to_x<B0>(c))) * 8u, r[e * R0 + c]);
                group_sync<GROUP>(grp);
#pragma unroll
                for (int y = 0; y < 16; ++y) r[y] = lds64(ga1 + uint32_t(C::pad(y << B0)) * 8u);
                radix16(r, tw1a, tw1b);
                if constexpr (C::NP == 3) {
#pragma unroll
                    for (int y = 0; y < 16; ++y) sts64(ga1 + uint32_t(C::pad(y << B0)) * 8u, r[y]);
                    __syncwarp();
#pragma unroll
                    for (int y = 0; y < 16; ++y) r[y] = lds64(ga2 + uint32_t(C::pad(y << (B0 + 4))) * 8u);
                    radix16(r, tw2a, tw2b);
                }

                RXPROF(3);   // FFT
                // ---- peak scan over this thread's bins tL + G*y (LoRaDetector.hpp:36-48) ----------------
                float m2[16];
#pragma unroll
                for (int y = 0; y < 16; ++y) m2[y] = pcx_mag2(r[y]);
                float bv = 0.0f; int by = 0;
                double sum = 0.0;
#pragma unroll
                for (int y = 0; y < 16; ++y) {
                    if (m2[y] > bv) { bv = m2[y]; by = y; }
                    sum += double(m2[y]);
                }
                int bi = tL + G * by;
                constexpr int WIDTH = G < 32 ? G : 32;
#pragma unroll
                for (int o = WIDTH / 2; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    argmax_merge(bv, bi, ov, oi);
                }
                if constexpr (G > 32) {
                    constexpr int KW = G / 32;
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
                                 : "r"(rbase + uint32_t(lane & (KW - 1)) * 16u));
                    sum = __longlong_as_double(ps);
                    bv = __uint_as_float(uint32_t(pk));
                    bi = int(uint32_t(pk >> 32));
#pragma unroll
                    for (int o = KW / 2; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        sum += __shfl_xor_sync(0xffffffffu, sum, o);
                        argmax_merge(bv, bi, ov, oi);
                    }
                }
                // neighbours of the peak live in other threads' registers
                const unsigned jl = (unsigned(bi) + unsigned(N - 1)) & unsigned(N - 1);
                const unsigned jr = (unsigned(bi) + 1u) & unsigned(N - 1);
                const int tl = int(jl & unsigned(G - 1)), yl = int(jl >> (SF - 4));
                const int tr = int(jr & unsigned(G - 1)), yr = int(jr >> (SF - 4));
                c64 nbv = r[0];
                if (G <= 32 || __any_sync(0xffffffffu, tL == tl || tL == tr)) nbv = pick16(r, tL == tl ? yl : yr);
                if (live) {
                    RawResult *slot = rawres + k;
                    if (tL == tl) slot->left = cunpack(nbv);
                    if (tL == tr) slot->right = cunpack(nbv);
                    if (t == 0) { slot->index = bi; slot->max_mag2 = bv; slot->total = sum; }
                }
                group_sync<GROUP>(grp);   // the record is complete; the exchange buffer is free again
                RXPROF(4);   // peak scan, reductions, neighbours
                if (live && t == 0) {
                    RxRoundRes o;
                    finalize_device(rawres[k], p.power_scale, &o.r, s_need_f != 0);
                    o.idx_after = walking ? s_after[k] : s_x0;
                    o.max_mag2 = rawres[k].max_mag2;
                    o.total = rawres[k].total;
                    res[k] = o;
                }
                RXPROF(5);   // finalize
            }
        };
        // thread 0 walks the round's results through the state machine while the speculation holds
        auto walk = [&]() {
                const int have = s_nwin;                        // windows demodulated in this round
                const unsigned long long n_avail = (p.n_samples - off) / N;
                int k = 0;
                s_force2 = 0;
                while (k < have && unsigned(npk) < p.max_packets) {
                    // the reference only works on a window while two symbols are buffered (LoRaDemod.cpp:148)
                    if ((unsigned long long)k + 2 > n_avail) break;
                    const lorab200_result r0 = res[k].r;
                    const bool need2 = rx_needs_second(p.cfg, st, r0);
                    if (need2 && k + 1 >= have) { s_force2 = 1; break; }   // look-ahead not demodulated yet: next round
                    if (st.phase == RX_DOWNCHIRP1) { st.sig_max_mag2 = res[k].max_mag2; st.sig_total = res[k].total; }
                    const RxAction a = rx_step(p.cfg, st, r0, res[k].idx_after, need2, need2 ? res[k + 1].r : r0);
                    if (a.emit_symbol >= 0 && st.sym_count >= 1 && st.sym_count <= p.cfg.mtu)
                        slots[size_t(npk) * p.cfg.mtu + (st.sym_count - 1)] = int16_t(a.emit_symbol);
                    if (a.post_packet) {
                        info[npk].length = st.sym_count;
                        info[npk].error = st.sig_error;
                        info[npk].power = st.sig_power;
                        info[npk].snr = st.sig_snr;
                        info[npk].max_mag2 = st.sig_max_mag2;
                        info[npk].reserved = 0.0f;
                        info[npk].total = st.sig_total;
                        npk++;
                    }
                    off += a.consume;
                    // is the next demodulated window the one the stream continues with?
                    // (... and, if the round skipped the fractional index, not a window that needs it)
                    const bool same = a.consume == uint32_t(N) && st.table == s_table &&
                                      __fmul_rn(st.ffe, 128.0f) == s_d && st.fine_index == res[k].idx_after &&
                                      (s_need_f != 0 || st.phase != RX_FRAMESYNC);
                    if (!same) { k = have; break; }
                    ++k;
                }
                if (k == 0 && !s_force2 && s_go) {
                    // no progress is only legal when the look-ahead was missing (s_force2); guard against a stall
                    s_force2 = 1;
                }
        };
        // One call site each for compute / walk / plan (the kernel is ~90 KB of code; a second inlined copy
        // for the one-warp rounds made many-channel launches fetch-bound).  A one-warp (solo) stretch: warp 0
        // iterates alone with warp-level syncs, the other warps wait at the barrier at the top of the loop.
        if (tid == 0) plan();
        for (;;) {
            __syncthreads();
            if (!s_go) break;
            const bool solo = s_solo != 0;
            if (solo && warp_u != 0) continue;
            bool again;
            do {
#ifdef LORAB200_RX_PROF
                if (tid == 0) {
                    prof_on = (s_solo || (LORAB200_RX_PROF > 1 && s_nwin <= 2)) ? 1 : 0;
                    asm volatile("mov.u64 %0, %%clock64;" : "=l"(prof_last) :: "memory");
                }
                if (solo) __syncwarp(); else __syncthreads();
#endif
                compute();
                if (solo) __syncwarp(); else __syncthreads();
                if (tid == 0) { walk(); RXPROF(6); plan(); RXPROF(7); }
#ifdef LORAB200_RX_PROF
                if (tid == 0 && prof_on) prof_rounds++;
#endif
                again = false;
                if (solo) { __syncwarp(); again = s_go && s_solo; }
            } while (again);
        }
        if (st.phase == RX_DATASYMBOLS && unsigned(npk) < p.max_packets)
            for (int k = tid; k < st.sym_count; k += K::CTA)
                p.partial[size_t(ch) * p.cfg.mtu + k] = slots[size_t(npk) * p.cfg.mtu + k];
        // slots that hold no packet read as empty (a consumer may scan all max_packets entries)
        for (unsigned q = unsigned(npk) + tid; q < p.max_packets; q += K::CTA) info[q].length = 0;
        if (tid == 0) { p.state[ch] = st; p.n_packets[ch] = npk; p.consumed[ch] = off; }
        __syncthreads();
    }
#ifdef LORAB200_RX_PROF
    if (threadIdx.x == 0 && prof_rounds > 0)
        printf("RXPROF sf=%d rounds=%lld top=%lld walkidx=%lld load=%lld fft=%lld reduce=%lld finalize=%lld walk=%lld plan=%lld\n", SF,
               prof_rounds, prof_acc[0], prof_acc[1], prof_acc[2], prof_acc[3], prof_acc[4], prof_acc[5], prof_acc[6], prof_acc[7]);
#endif
}

} // namespace lorab200
