// tables.cpp -- host-side table recipes of the demod hot path (product code).
//
// The GPU kernels reproduce the reference's float dataflow bit-for-bit, so the
// tables they multiply by must hold exactly the values the reference computes
// with the host libm.  They are built once per context on the host and
// uploaded.
//
// Reference recipes (file:line relative to myriadrf/LoRa-SDR):
//   dechirp tables   LoRaDemod.cpp:97-107
//   fine-tune table  LoRaDemod.cpp:108-114
//   FFT twiddles     kissfft.hh:17-22
//   genChirp         ChirpGenerator.hpp:21-47
#include "tables.h"

#include <cmath>
#include <complex>

namespace lorab200 {

// Conjugate-up and down dechirp tables.  The reference advances a *float*
// instantaneous frequency by 2*pi/N per sample (rounded to float every step)
// and integrates it in a *double* phase; each entry is polar(1.0, phase) in
// double, narrowed to float (the "up" table stores its conjugate).
void build_dechirp_tables(int sf, std::vector<cfloat> &up, std::vector<cfloat> &down)
{
    const size_t n = size_t(1) << sf;
    up.resize(n);
    down.resize(n);
    float freq = float(-M_PI);
    double angle = 0.0;
    const double dfreq = (2 * M_PI) / double(n);
    for (size_t i = 0; i < n; ++i) {
        angle += double(freq);
        const std::complex<double> e = std::polar(1.0, angle);
        up[i] = cfloat{float(e.real()), float(-e.imag())};
        down[i] = cfloat{float(e.real()), float(e.imag())};
        freq = float(double(freq) + dfreq);
    }
}

// Fine-tune phasors: entry i = polar(1, (i+1)*step) with the step held in a
// float and the running angle in a double.  Entry 0 is therefore NOT 1+0j.
void build_fine_table(int sf, std::vector<cfloat> &fine)
{
    const size_t n = size_t(1) << sf;
    const size_t len = n * kFineSteps;
    fine.resize(len);
    const float step = float(2.0 * M_PI / double(len));
    double angle = 0.0;
    for (size_t i = 0; i < len; ++i) {
        angle += double(step);
        const std::complex<double> e = std::polar(1.0, angle);
        fine[i] = cfloat{float(e.real()), float(e.imag())};
    }
}

// kissfft forward twiddles: tw[i] = exp(j * float(i) * phinc), with
// phinc = -2*acosf(-1)/n evaluated in float; std::exp(std::complex<float>)
// is libm's cexpf, so this is bit-identical to the reference on any host.
void build_twiddles(int sf, std::vector<cfloat> &tw)
{
    const int n = 1 << sf;
    tw.resize(size_t(n));
    const float phinc = -2.0f * std::acos(-1.0f) / float(n);
    for (int i = 0; i < n; ++i) {
        const std::complex<float> e = std::exp(std::complex<float>(0.0f, float(i) * phinc));
        tw[size_t(i)] = cfloat{e.real(), e.imag()};
    }
}

// 20*log10(N) as LoRaDetector.hpp:18 evaluates it (double log10, stored float).
float power_scale(int sf)
{
    return float(20 * std::log10(double(size_t(1) << sf)));
}

// genChirp<float>: all-float frequency and phase accumulators, frequency
// incremented before use, wrap of f at fMax, phase reduced mod 2*pi at the end.
int gen_chirp(cfloat *samps, int N, int ovs, int NN, float f0, bool down, float ampl,
              float &phase_accum)
{
    const float f_lo = float(-M_PI / ovs);
    const float f_hi = float(M_PI / ovs);
    const float f_inc = float((2 * M_PI) / (N * ovs * ovs));
    const float span = f_hi - f_lo;
    float f = f_lo + f0;
    float ph = phase_accum;
    for (int i = 0; i < NN; ++i) {
        f += f_inc;
        if (f > f_hi) f -= span;
        ph = down ? ph - f : ph + f;
        const std::complex<float> s = std::polar(ampl, ph);
        samps[i] = cfloat{s.real(), s.imag()};
    }
    ph = float(double(ph) - std::floor(double(ph) / (2 * M_PI)) * 2 * M_PI);
    phase_accum = ph;
    return NN > 0 ? NN : 0;
}

} // namespace lorab200
