// fine_orbit.h -- the demod's fine-tune index (LoRaDemod.cpp:160-162) at ANY sample of a batch in
// O(1), as a kernel parameter: no prefix kernels, no per-chunk state array.
//
// The index is advanced once per sample by g(x) = wrap(trunc(fl32(float(x) - d))).  fine_walk.h shows
// that, away from the wrap of the table and unless d sits within 2^-5 of an integer (a finite Kp),
// every step subtracts the same integer s (s = d for an integer d, floor(d) + 1 otherwise): the walk is
// a sequence of arithmetic progressions ("segments"), glued by a few literal steps around each wrap.
// After the first wrap the states at which a segment starts can only take a handful of values near the
// end of the table, so the sequence of segments becomes periodic within a few wraps.  FineOrbit stores
// the segments before the period and one period; the state at sample n is then
//
//     n' = n                                   (n  < pre_len)
//     n' = pre_len + (n - pre_len) mod period  (n >= pre_len)
//     X(n) = seg[j].y - (n' - seg[j].start) * s       for the segment j that contains n'.
//
// The builder only ever uses FineWalk::step1 (the literal recurrence) to cross from one segment to the
// next and checks the first and last transition of every progression against it, so the table is exact
// by construction; tests/test_fine_walk.py compares orbit_at() with the sequential walk.  Steps that do
// not fit (finite Kp, s == 0, more than kMaxSeg segments before the period closes) are reported as "no
// orbit": the caller falls back to the prefix path (fine_prefix.cuh).
#pragma once
#include <cstdint>

#include "fine_walk.h"

#if defined(__CUDACC__)
#define LB_ORBIT_HD __host__ __device__ __forceinline__
#else
#define LB_ORBIT_HD inline
#endif

namespace lorab200 {

constexpr int kMaxOrbitSeg = 48;

struct FineOrbitSeg {
    long long start; // sample offset (from the start of the batch) of the segment's first state
    int y;           // state at `start`
    int c;           // the states at start + r, r = 0..c, are y - r*s
};

struct FineOrbit {
    int n_seg;            // segments in seg[]: [0, n_pre) before the period, [n_pre, n_seg) one period
    int n_pre;
    int s;                // the common difference
    int pad_;
    long long pre_len;    // samples covered by the segments before the period
    long long period;     // samples covered by one period
    FineOrbitSeg seg[kMaxOrbitSeg];
};

// (j, r) with r possibly beyond segment j  ->  the segment / offset that holds that sample, and its state
LB_ORBIT_HD int orbit_state(const FineOrbit &o, int &j, int &r)
{
    while (r > o.seg[j].c) {
        r -= o.seg[j].c + 1;
        ++j;
        if (j == o.n_seg) j = o.n_pre;
    }
    return o.seg[j].y - r * o.s;
}

// segment / offset of sample n (64-bit modulo + a scan of the segment starts: once per thread, then
// orbit_state() advances incrementally)
LB_ORBIT_HD void orbit_locate(const FineOrbit &o, unsigned long long n, int &j, int &r)
{
    long long m = (long long)n;
    int lo = 0, hi = o.n_pre;
    if (m >= o.pre_len) {
        m = o.pre_len + (long long)((unsigned long long)(m - o.pre_len) % (unsigned long long)o.period);
        lo = o.n_pre;
        hi = o.n_seg;
    }
    j = lo;
    for (int q = lo + 1; q < hi; ++q)
        if (o.seg[q].start <= m) j = q;
    r = int(m - o.seg[j].start);
}

LB_ORBIT_HD int orbit_at(const FineOrbit &o, unsigned long long n)
{
    int j, r;
    orbit_locate(o, n, j, r);
    return orbit_state(o, j, r);
}

// Host: build the orbit of x0 under the recurrence with step d on a table of `len` entries.
// Returns false when the walk has no such description (see the header comment).
inline bool build_fine_orbit(const FineWalk &fw, int32_t x0, FineOrbit &o)
{
    const int64_t none = INT64_MAX;
    if (!fw.integer && fw.Kp != none) return false;
    const int64_t s = fw.integer ? fw.D : fw.D + 1;
    if (s == 0 || s >= int64_t(fw.len) || -s >= int64_t(fw.len)) return false;
    // a transition from x is "linear" when it is x -> x - s without a wrap
    auto linear_run = [&](int64_t y) -> int64_t { // number of consecutive linear transitions starting at y
        if (s > 0) {
            const int64_t lo = fw.integer ? s : (fw.D + 1 > 0 ? fw.D + 1 : 0); // smallest linear state
            return y >= lo ? (y - lo) / s + 1 : 0;
        }
        const int64_t hi = int64_t(fw.len) - 1 + s;                            // largest linear state (s < 0)
        const int64_t lo = fw.integer ? 0 : (fw.D + 1 > 0 ? fw.D + 1 : 0);     // k >= 1
        if (y < lo || y > hi) return 0;
        return (hi - y) / (-s) + 1;
    };
    o = FineOrbit{};
    o.s = int(s);
    int64_t y = x0, start = 0;
    for (int i = 0; i < kMaxOrbitSeg + 1; ++i) {
        // does a segment starting in state y exist already?  then the period closes here
        for (int k = 0; k < i; ++k) {
            if (o.seg[k].y == int(y)) {
                o.n_pre = k;
                o.n_seg = i;
                o.pre_len = o.seg[k].start;
                o.period = start - o.seg[k].start;
                return o.period > 0;
            }
        }
        if (i == kMaxOrbitSeg) return false;
        const int64_t c = linear_run(y);
        // the progression must agree with the literal recurrence at both ends
        if (c > 0) {
            if (FineWalk::step1(int32_t(y), fw.d, fw.len) != int32_t(y - s)) return false;
            const int64_t last = y - (c - 1) * s;
            if (FineWalk::step1(int32_t(last), fw.d, fw.len) != int32_t(last - s)) return false;
        }
        o.seg[i].start = start;
        o.seg[i].y = int(y);
        o.seg[i].c = int(c);
        const int64_t tail = y - c * s;                       // last state of the segment
        y = FineWalk::step1(int32_t(tail), fw.d, fw.len);     // literal step into the next segment
        start += c + 1;
    }
    return false;
}

} // namespace lorab200
