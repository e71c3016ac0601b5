// fine_walk.h -- closed-form advance of the demod's fine-tune index recurrence.
//
// LoRaDemod.cpp:160-162 advances an int index once per sample with float arithmetic:
//
//     _fineTuneIndex -= _finefreqError * _fineSteps;        // int = int(float(int) - float)
//     if (_fineTuneIndex < 0) _fineTuneIndex += N*_fineSteps;
//     else if (_fineTuneIndex >= int(N*_fineSteps)) _fineTuneIndex -= N*_fineSteps;
//
// i.e. x <- wrap(trunc(fl32(float(x) - d))).  A batched demodulator needs the index at the
// start of every window (N steps apart) without walking W*N samples on the host.
//
// Structure of one step.  Write d = D + phi with D = floor(d), phi in [0,1), and k = x - D.
// For phi == 0 the step is exactly x - D.  Otherwise the exact difference u = k - phi lies
// strictly between the integers k-1 and k:
//   * k >= 1: trunc(fl32(u)) is k-1 unless the float rounding carries u up to k, which
//     depends only on the float spacing just below k -- it is monotone in k: there is a
//     threshold Kp with  step = D+1 for 1 <= k < Kp  and  step = D for k >= Kp
//     (for phi >= 2^-5 rounding never carries: Kp = infinity);
//   * k <= 0: (a handful of states per wrap of the table) is evaluated directly.
// Inside a region the step is a constant, so j steps are one multiplication.  Every region
// boundary, and everything that is not provably constant, falls back to evaluating the
// recurrence itself, so the result is exact by construction; tests/test_fine_walk.py checks
// it against the sequential recurrence for many (x, d, n), including the awkward d.
#pragma once
#include <cstdint>

namespace lorab200 {

struct FineWalk {
    int32_t len;     // table length N*128
    float   d;       // step per sample (_finefreqError * _fineSteps)
    bool    integer; // phi == 0
    int64_t D;       // floor(d)
    int64_t Kp;      // smallest k >= 1 whose step is D (rounding carries); INT64_MAX if none

    // one step, exactly as the reference evaluates it
    static inline int32_t step1(int32_t x, float d, int32_t len)
    {
        int32_t y = int32_t(float(x) - d);
        if (y < 0) y += len;
        else if (y >= len) y -= len;
        return y;
    }
    // raw truncated value for a hypothetical index x (no wrap), used to locate Kp
    static inline int64_t raw(int64_t x, float d) { return int64_t(float(x) - d); }

    FineWalk(int32_t len_, float d_);
    int32_t advance(int32_t x, uint64_t n) const; // g^n(x)
};

} // namespace lorab200
