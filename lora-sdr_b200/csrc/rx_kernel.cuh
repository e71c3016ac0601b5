// rx_kernel.cuh -- sync acquisition and symbol extraction on the device (row f3): one CTA per
// channel runs the reference demod's whole work() sequence over its stream -- demodulate the
// window at the current offset (the simple kernel's cooperative window routine), advance the
// frame-sync state machine (rx_fsm.h, the restatement of LoRaDemod.cpp:173-327), move the
// offset by what the state machine consumed, repeat -- and writes the symbol packets the
// reference block would have posted.  Hundreds of channels run side by side in one launch; no
// host round trip per window, which is what made a single not-squelched stream slower on the
// GPU than on the CPU at the block boundary.
#pragma once
#include "demod_simple.cuh"
#include "fine_prefix.cuh"
#include "rx_fsm.h"

namespace lorab200 {

struct RxPacketInfo {
    int32_t length;     // symbols in the packet
    int32_t error;      // the "error" signal (LoRaDemod.cpp:267)
    float power, snr;   // the "power" / "snr" signals (:268-269)
    float max_mag2, reserved;   // raw detector outputs behind them
    double total;
};

struct RxParams {
    const float2 *iq;                 // channel c at iq + c * chan_stride
    unsigned long long chan_stride;   // samples
    unsigned long long n_samples;     // per channel
    unsigned int n_channels;
    const float2 *up, *down, *fine, *tw;
    float power_scale;
    int fine_len;
    RxCfg cfg;
    RxState *state;                   // [n_channels], in/out (all zero = activate())
    int16_t *partial;                 // [n_channels][mtu], in/out: the packet under construction
    int16_t *symbols;                 // [n_channels][max_packets][mtu]
    RxPacketInfo *info;               // [n_channels][max_packets]
    int32_t *n_packets;               // [n_channels]
    unsigned long long *consumed;     // [n_channels]
    unsigned int max_packets;
};

// FineWalk's constructor (fine_walk.cpp) on the device: D = floor(d), the integer flag and the
// smallest k >= 1 whose step is D (binary search over the monotone "rounding carries" predicate)
__device__ inline FineWalkDev fine_walk_make_dev(int len, float d)
{
    FineWalkDev fw;
    fw.len = len; fw.d = d;
    const double dd = double(d), fl = floor(dd);
    fw.D = (long long)fl;
    fw.integer = (dd == fl) ? 1 : 0;
    fw.Kp = 0x7fffffffffffffffll;
    if (!fw.integer) {
        auto carries = [&](long long k) {
            return __float2ll_rz(__fsub_rn(__ll2float_rn(k + fw.D), d)) == k;
        };
        const long long kmax = (long long)len + (fw.D < 0 ? -fw.D : fw.D) + 2;
        if (carries(kmax)) {
            long long lo = 1, hi = kmax;
            while (lo < hi) {
                const long long mid = lo + (hi - lo) / 2;
                if (carries(mid)) hi = mid; else lo = mid + 1;
            }
            fw.Kp = lo;
        }
    }
    return fw;
}

template <int SF>
__global__ void __launch_bounds__(kSimpleThreads) rx_kernel(const RxParams p)
{
    constexpr int N = 1 << SF;
    constexpr int T = kSimpleThreads;
    constexpr int C = N / T;                                   // fine-tune chunk per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *F = reinterpret_cast<float2 *>(smem_raw);
    int *fidx = reinterpret_cast<int *>(F + N);
    __shared__ float s_val[T / 32];
    __shared__ int s_idx[T / 32];
    __shared__ double s_sum[T / 32];
    __shared__ RxState st;
    __shared__ FineWalkDev fw;
    __shared__ lorab200_result res[2];
    __shared__ float raw_max0;
    __shared__ double raw_total0;
    __shared__ unsigned long long off;
    __shared__ int go, need2, idx_after[2], walk_now, npk;

    const int tid = threadIdx.x;
    for (unsigned ch = blockIdx.x; ch < p.n_channels; ch += gridDim.x) {
        const float2 *in = p.iq + (unsigned long long)ch * p.chan_stride;
        int16_t *slots = p.symbols + (size_t(ch) * p.max_packets) * p.cfg.mtu;
        RxPacketInfo *info = p.info + size_t(ch) * p.max_packets;
        if (tid == 0) { st = p.state[ch]; off = 0; npk = 0; }
        __syncthreads();
        // a packet was under construction when the previous call ran out of samples
        if (st.phase == RX_DATASYMBOLS && p.max_packets > 0)
            for (int k = tid; k < st.sym_count; k += T) slots[k] = p.partial[size_t(ch) * p.cfg.mtu + k];
        __syncthreads();

        for (;;) {
            if (tid == 0) go = (p.n_samples - off >= 2ull * N) && (unsigned(npk) < p.max_packets);
            __syncthreads();
            if (!go) break;
            const float2 *chirp = st.table == LORAB200_TABLE_DOWN ? p.down : p.up;
            for (int w = 0; w < 2; ++w) {
                // fine-tune indices of this window: start index and step (LoRaDemod.cpp:159-162, :192-197)
                const int x0 = (w == 0) ? st.fine_index : idx_after[0];
                const float d = __fmul_rn(st.ffe, 128.0f);        // _finefreqError * _fineSteps
                if (tid == 0) { walk_now = (d != 0.0f); if (walk_now) fw = fine_walk_make_dev(p.fine_len, d); }
                __syncthreads();
                if (walk_now) {
                    int x = fine_advance_dev(fw, x0, (unsigned long long)(tid * C));
#pragma unroll 4
                    for (int i = 0; i < C; ++i) { fidx[tid * C + i] = x; x = fine_next(x, d, p.fine_len); }
                    if (tid == T - 1) idx_after[w] = x;    // after the window's N samples
                } else if (tid == 0) {
                    idx_after[w] = x0;
                }
                __syncthreads();
                const RawResult raw = simple_window<SF>(in + off + size_t(w) * N, chirp, p.fine, fidx, x0, walk_now != 0,
                                                        p.tw, nullptr, nullptr, F, s_val, s_idx, s_sum);
                if (tid == 0) {
                    finalize_device(raw, p.power_scale, &res[w]);
                    if (w == 0) { need2 = rx_needs_second(p.cfg, st, res[0]) ? 1 : 0; raw_max0 = raw.max_mag2; raw_total0 = raw.total; }
                }
                __syncthreads();
                if (w == 0 && !need2) break;
            }
            if (tid == 0) {
                if (st.phase == RX_DOWNCHIRP1) { st.sig_max_mag2 = raw_max0; st.sig_total = raw_total0; }
                const RxAction a = rx_step(p.cfg, st, res[0], idx_after[0], need2 != 0, res[1]);
                if (a.emit_symbol >= 0 && st.sym_count >= 1 && st.sym_count <= p.cfg.mtu)   // (sym_count already incremented)
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
            }
            __syncthreads();
        }
        // park the packet under construction for the next call, write the state back
        if (st.phase == RX_DATASYMBOLS && unsigned(npk) < p.max_packets)
            for (int k = tid; k < st.sym_count; k += T) p.partial[size_t(ch) * p.cfg.mtu + k] = slots[size_t(npk) * p.cfg.mtu + k];
        // slots that hold no packet read as empty (a consumer may scan all max_packets entries)
        for (unsigned q = unsigned(npk) + tid; q < p.max_packets; q += T) info[q].length = 0;
        if (tid == 0) { p.state[ch] = st; p.n_packets[ch] = npk; p.consumed[ch] = off; }
        __syncthreads();
    }
}

} // namespace lorab200
