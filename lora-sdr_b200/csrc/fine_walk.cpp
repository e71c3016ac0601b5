// fine_walk.cpp -- see fine_walk.h.  Host code, compiled with -ffp-contract=off and no
// -march so that float(x) - d is the same single-precision subtraction the reference performs.
#include "fine_walk.h"

#include <cmath>
#include <limits>

namespace lorab200 {

FineWalk::FineWalk(int32_t len_, float d_) : len(len_), d(d_)
{
    const double dd = double(d);
    const double fl = std::floor(dd);
    D = int64_t(fl);
    integer = (dd == fl);
    Kp = std::numeric_limits<int64_t>::max();
    if (!integer) {
        // carries(k): does fl32(float(k + D) - d) round up to the integer k?  Monotone in k >= 1.
        auto carries = [&](int64_t k) { return raw(k + D, d) == k; };
        const int64_t kmax = int64_t(len) + (D < 0 ? -D : D) + 2;
        if (carries(kmax)) {
            int64_t lo = 1, hi = kmax; // smallest k with carries(k)
            while (lo < hi) {
                const int64_t mid = lo + (hi - lo) / 2;
                if (carries(mid)) hi = mid; else lo = mid + 1;
            }
            Kp = lo;
        }
    }
}

int32_t FineWalk::advance(int32_t x, uint64_t n) const
{
    if (n == 0) return x;
    if (integer) {
        // exact integer step with a single wrap per sample == modular arithmetic (|D| < len)
        const int64_t L = len;
        // (n mod L) * (D mod L) is below L^2 < 2^40: fits an int64
        int64_t m = (int64_t(x) - int64_t(n % uint64_t(L)) * (D % L)) % L;
        if (m < 0) m += L;
        return int32_t(m);
    }
    while (n > 0) {
        const int64_t k = int64_t(x) - D;
        uint64_t j = 0;
        int64_t s = 0;
        if (k >= 1) {
            const bool r = (k >= Kp);
            s = D + 1 - (r ? 1 : 0);
            if (s == 0) return x; // fixed point: u = x - phi' truncates back to x
            if (s > 0) {
                const int64_t lb = r ? Kp : 1;                 // pre-states k' >= lb keep the same step
                j = uint64_t((k - lb) / s) + 1;
            } else {
                // x grows by |s| per step; stop before the wrap (y >= len) and, for r == 0,
                // before k reaches Kp
                const int64_t a = -s;
                int64_t room = (int64_t(len) - 1 - int64_t(x)) / a;   // steps with y <= len-1
                if (!r && Kp != std::numeric_limits<int64_t>::max()) {
                    const int64_t q = (Kp - 1 - k) / a + 1;           // pre-states k' <= Kp-1
                    if (q < room) room = q;
                }
                j = room > 0 ? uint64_t(room) : 0;
            }
        }
        if (j == 0) {
            // region boundary, wrap, or k <= 0: evaluate the recurrence itself
            const int32_t y = step1(x, d, len);
            if (y == x) return x; // stuck (e.g. 0 < d < 1 parks the index at 0)
            x = y;
            --n;
            continue;
        }
        if (j > n) j = n;
        x = int32_t(int64_t(x) - int64_t(j) * s);
        n -= j;
    }
    return x;
}

} // namespace lorab200
