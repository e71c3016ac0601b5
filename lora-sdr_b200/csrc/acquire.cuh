// acquire.cuh -- preamble search over ALL candidate offsets (SURVEY.md section 8 row f3 as written:
// "speculatively demodulate at all candidate offsets (stride < N) to remove the sequential N - value
// re-timing").  Behavioural, not bit, parity: the reference finds a preamble by hunting for it one window
// at a time (LoRaDemod.cpp:179-237: every un-squelched window consumes N - value samples to re-time the
// stream, :219); here the fused kernel demodulates the windows at offsets 0, S, 2S, ... (S = stride, any
// even number of samples, typically N/4) in one launch and three small kernels turn the result records into
// a list of preambles, in stream order.
//
// Inside a train of identical up-chirps a window that starts delta samples after a chirp boundary dechirps
// to a tone at bin delta (that is what `total = N - value` relies on), so the quantity
//     phase(w) = (offset(w) - index(w)) mod N
// is the same for every window of the train, whatever its offset; a carrier offset adds a constant.  A
// PREAMBLE is a maximal run of consecutive windows with snr = power - power_avg >= thresh (the reference's
// squelch test, LoRaDemod.cpp:173-174) whose phase moves by at most one bin from window to window, at least
// min_windows long.  For every run the record carries the first window's offset, the phase, the offset the
// reference's re-timing would jump to from that window (offset + N - index, :219), the run length and the
// mean snr / peak power -- everything a receiver needs to start its state machine there.
#pragma once
#include "common.cuh"

namespace lorab200 {

constexpr int kAcqTile = 1024;            // windows per CTA (one per thread)
constexpr int kAcqMaxRun = 1 << 16;       // a run is measured up to this many windows

struct AcqParams {
    const lorab200_result *res;           // [n_win] records of the windows at offset w * stride
    unsigned long long n_win, stride;
    int n;                                // N = 2^SF
    float thresh;
    int min_windows;
    unsigned char *valid;                 // [n_win] scratch: 1 at the first window of a reported run
    int *tile_count, *tile_base;          // [n_tiles] scratch
    unsigned int n_tiles;
    lorab200_preamble *found;
    int *n_found;
    unsigned int max_found;
};

__device__ __forceinline__ bool acq_cand(const lorab200_result &r, float thresh)
{
    return (r.power - r.power_avg) >= thresh;     // false for NaN, like the reference's `snr < thresh` is
}
__device__ __forceinline__ int acq_phase(const AcqParams &p, unsigned long long w, int index)
{
    const int o = int((w * p.stride) & (unsigned long long)(p.n - 1));
    return (o - index) & (p.n - 1);
}
__device__ __forceinline__ bool acq_close(int a, int b, int n)
{
    const int d = (a - b) & (n - 1);
    return d <= 1 || d == n - 1;
}
// is w the first window of a run?
__device__ __forceinline__ bool acq_is_start(const AcqParams &p, unsigned long long w, const lorab200_result &r)
{
    if (!acq_cand(r, p.thresh)) return false;
    if (w == 0) return true;
    const lorab200_result q = p.res[w - 1];
    return !(acq_cand(q, p.thresh) && acq_close(acq_phase(p, w - 1, q.index), acq_phase(p, w, r.index), p.n));
}
// windows in the run that starts at w (up to kAcqMaxRun), optionally its statistics
__device__ __forceinline__ int acq_run(const AcqParams &p, unsigned long long w, lorab200_result r, float *snr_sum, float *pmax)
{
    int len = 1;
    float ss = r.power - r.power_avg, pm = r.power;
    int ph = acq_phase(p, w, r.index);
    while (len < kAcqMaxRun && w + len < p.n_win) {
        const lorab200_result q = p.res[w + len];
        const int pq = acq_phase(p, w + len, q.index);
        if (!acq_cand(q, p.thresh) || !acq_close(ph, pq, p.n)) break;
        ss += q.power - q.power_avg;
        pm = fmaxf(pm, q.power);
        ph = pq;
        ++len;
    }
    if (snr_sum) *snr_sum = ss;
    if (pmax) *pmax = pm;
    return len;
}

// one thread per window: mark the starts of runs that are long enough, count them per tile
__global__ void __launch_bounds__(kAcqTile) acq_mark_kernel(const AcqParams p)
{
    for (unsigned int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const unsigned long long w = (unsigned long long)tile * kAcqTile + threadIdx.x;
        int v = 0;
        if (w < p.n_win) {
            const lorab200_result r = p.res[w];
            if (acq_is_start(p, w, r)) v = acq_run(p, w, r, nullptr, nullptr) >= p.min_windows ? 1 : 0;
            p.valid[w] = (unsigned char)v;
        }
        const int cnt = __syncthreads_count(v);
        if (threadIdx.x == 0) p.tile_count[tile] = cnt;
    }
}

// exclusive scan of the tile counts (one CTA; a few thousand tiles at most per call)
__global__ void __launch_bounds__(1024) acq_scan_kernel(const AcqParams p)
{
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned int base = 0; base < p.n_tiles; base += 1024) {
        const unsigned int t = base + threadIdx.x;
        const int c = t < p.n_tiles ? p.tile_count[t] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int s = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += up;
            }
            warp_sum[lane] = s;     // inclusive over warps
        }
        __syncthreads();
        const int before = carry + (wid ? warp_sum[wid - 1] : 0) + incl - c;
        if (t < p.n_tiles) p.tile_base[t] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) *p.n_found = carry;
}

// one thread per window again: the marked starts write their records at their rank in stream order
__global__ void __launch_bounds__(kAcqTile) acq_emit_kernel(const AcqParams p)
{
    __shared__ int warp_sum[kAcqTile / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        if (p.tile_count[tile] == 0) continue;                       // (uniform per CTA)
        const unsigned long long w = (unsigned long long)tile * kAcqTile + threadIdx.x;
        const int v = (w < p.n_win) ? int(p.valid[w]) : 0;
        const unsigned int ballot = __ballot_sync(0xffffffffu, v != 0);
        if (lane == 0) warp_sum[wid] = __popc(ballot);
        __syncthreads();
        int before = p.tile_base[tile] + __popc(ballot & ((1u << lane) - 1u));
        for (int q = 0; q < wid; ++q) before += warp_sum[q];
        if (v && unsigned(before) < p.max_found) {
            const lorab200_result r = p.res[w];
            float ss, pm;
            const int len = acq_run(p, w, r, &ss, &pm);
            lorab200_preamble o;
            o.offset = w * p.stride;
            o.aligned = o.offset + (unsigned long long)(p.n - r.index);
            o.phase = acq_phase(p, w, r.index);
            o.windows = len;
            o.snr_db = ss / float(len);
            o.power_db = pm;
            p.found[before] = o;
        }
        __syncthreads();
    }
}

} // namespace lorab200
