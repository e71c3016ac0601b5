// finalize.cpp -- host-side completion of the detector outputs with the HOST libm.
//
// The kernels hand back, per window, the raw quantities LoRaDetector::detect has
// after its scan loop (peak index, peak |bin|^2, double total, the two
// neighbouring bins).  The dB conversions and the fractional index
// (LoRaDetector.hpp:50-61) need log10f / hypotf, whose glibc results are not
// reproducible bit-for-bit on the device, so the host-buffer entry point
// (lorab200_demod_host, the call a Pothos work() makes) finishes them here --
// O(1) flops per window -- and is bit-identical to the reference.
#include "finalize.h"

#include <cmath>
#include <complex>

namespace lorab200 {

void finalize_host(const RawResultHost *raw, size_t n, float power_scale, lorab200_result *out)
{
    for (size_t w = 0; w < n; ++w) {
        const RawResultHost &r = raw[w];
        const float peak2 = r.max_mag2;
        const float noise = std::sqrt(float(r.total - peak2));
        const float fundamental = std::sqrt(peak2);
        lorab200_result o;
        o.index = r.index;
        o.power_avg = 20 * std::log10(noise) - power_scale;
        o.power = 20 * std::log10(fundamental) - power_scale;
        const float left = std::abs(std::complex<float>(r.left_re, r.left_im));
        const float right = std::abs(std::complex<float>(r.right_re, r.right_im));
        const double denom = (2.0 * fundamental) - right - left;
        o.f_index = (denom == 0.0) ? 0.0f : float(0.5 * (right - left) / denom);
        out[w] = o;
    }
}

} // namespace lorab200
