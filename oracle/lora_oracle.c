/*
 * lora_oracle.c -- CPU restatement of the LoRa-SDR demod hot path
 * (dechirp -> kissfft -> peak detect).  TEST INFRASTRUCTURE ONLY; see the
 * header of lora_oracle.h for who may use it and for the parity-pinning
 * status (PINNED against the reference headers compiled in place).
 *
 * Written from the reference's behaviour, not copied: kissfft's recursive
 * decimation-in-time transform is restated as an explicit digit-reversed load
 * followed by iterative in-place stages (same operand graph, hence the same
 * float roundings; storage order is irrelevant to rounding).
 *
 * All file:line citations are relative to /root/reference.
 */
#define _GNU_SOURCE
#include "lora_oracle.h"

#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* std::complex<float> operator*: naive 4-mul form, every product and sum
 * rounded separately (GCC expands it like this for finite operands; the
 * __mulsc3 NaN-recovery path is out of contract).  Bitwise commutative. */
static inline lo_cf cmul(lo_cf a, lo_cf b)
{
    lo_cf r;
    float ac = a.re * b.re, bd = a.im * b.im;
    float ad = a.re * b.im, bc = a.im * b.re;
    r.re = ac - bd;
    r.im = ad + bc;
    return r;
}
static inline lo_cf cadd(lo_cf a, lo_cf b) { lo_cf r = { a.re + b.re, a.im + b.im }; return r; }
static inline lo_cf csub(lo_cf a, lo_cf b) { lo_cf r = { a.re - b.re, a.im - b.im }; return r; }

/* kissfft.hh:17-22  fill_twiddles, forward (inverse=false):
 *   T phinc = -2 * acos((T)-1) / nfft;  dst[i] = exp(complex<T>(0, i*phinc)).
 * std::exp(complex<float>) is __builtin_cexpf in libstdc++, i.e. libm's cexpf;
 * calling the same libm entry keeps the table bit-identical on any host. */
void lo_twiddles(int n, lo_cf *tw)
{
    float phinc = (float)(-2) * acosf((float)-1) / (float)n;
    for (int i = 0; i < n; ++i) {
        float complex e = cexpf(CMPLXF(0.0f, (float)i * phinc));
        tw[i].re = crealf(e);
        tw[i].im = cimagf(e);
    }
}

lo_plan *lo_plan_create(int sf)
{
    if (sf < 1 || sf > 24) return NULL;
    lo_plan *p = (lo_plan *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    p->sf = sf;
    p->n = 1 << sf;
    /* kissfft.hh:34-51  factor out 4s, then 2s (n is a power of two). */
    int n = p->n, radix = 4;
    do {
        while (n % radix) radix = 2;
        n /= radix;
        p->radix[p->nstages] = radix;
        p->remainder[p->nstages] = n;
        p->nstages++;
    } while (n > 1);
    p->twiddles = (lo_cf *)malloc(sizeof(lo_cf) * (size_t)p->n);
    lo_twiddles(p->n, p->twiddles);
    /* LoRaDetector.hpp:18  _powerScale = 20*std::log10(N): size_t -> double log10,
     * narrowed to the float member. */
    p->power_scale = (float)(20 * log10((double)p->n));
    /* kissfft.hh:90-94,101-103: output slot j (digits q_s = (j / m_s) % r_s) is
     * loaded from in[i], i = q_0 + r_0*(q_1 + r_1*(q_2 + ...)): every recursion
     * level decimates the input by its radix. */
    p->rev = (int *)malloc(sizeof(int) * (size_t)p->n);
    for (int j = 0; j < p->n; ++j) {
        int i = 0, w = 1;
        for (int s = 0; s < p->nstages; ++s) {
            int q = (j / p->remainder[s]) % p->radix[s];
            i += q * w;
            w *= p->radix[s];
        }
        p->rev[j] = i;
    }
    return p;
}

void lo_plan_destroy(lo_plan *p)
{
    if (!p) return;
    free(p->twiddles);
    free(p->rev);
    free(p);
}

/* LoRaDemod.cpp:97-107. float `phase` accumulates 2pi/N (rounded to float each
 * step), double phaseAccum, std::polar(1.0, double) = (cos, sin) in double. */
void lo_tables_demod(int sf, lo_cf *up, lo_cf *down)
{
    const size_t N = (size_t)1 << sf;
    float phase = (float)(-M_PI);
    double phaseAccum = 0.0;
    for (size_t i = 0; i < N; i++) {
        phaseAccum += phase;
        double er = 1.0 * cos(phaseAccum), ei = 1.0 * sin(phaseAccum);
        if (up)   { up[i].re = (float)er;   up[i].im = (float)(-ei); } /* conj(entry) */
        if (down) { down[i].re = (float)er; down[i].im = (float)ei; }
        phase += (2 * M_PI) / N; /* float = float(double(phase) + 2pi/N) */
    }
}

/* LoRaDemod.cpp:108-114. */
void lo_table_fine(int sf, lo_cf *fine)
{
    const size_t N = (size_t)1 << sf;
    const size_t fineSteps = 128;
    double phaseAccum = 0.0;
    float phase = 2.0 * M_PI / (N * fineSteps); /* `phase` is still the float variable */
    for (size_t i = 0; i < N * fineSteps; i++) {
        phaseAccum += phase;
        fine[i].re = (float)(1.0 * cos(phaseAccum));
        fine[i].im = (float)(1.0 * sin(phaseAccum));
    }
}

/* ChirpGenerator.hpp:21-47, Type=float. std::polar(float,float) =
 * (ampl*cosf(theta), ampl*sinf(theta)). */
int lo_gen_chirp(lo_cf *samps, int N, int ovs, int NN, float f0, int down,
                 float ampl, float *phaseAccum_io)
{
    const float fMin = -M_PI / ovs;
    const float fMax = M_PI / ovs;
    const float fStep = (2 * M_PI) / (N * ovs * ovs);
    float f = fMin + f0;
    float phaseAccum = *phaseAccum_io;
    int i;
    for (i = 0; i < NN; i++) {
        f += fStep;
        if (f > fMax) f -= (fMax - fMin);
        if (down) phaseAccum -= f; else phaseAccum += f;
        samps[i].re = ampl * cosf(phaseAccum);
        samps[i].im = ampl * sinf(phaseAccum);
    }
    phaseAccum -= floor(phaseAccum / (2 * M_PI)) * 2 * M_PI;
    *phaseAccum_io = phaseAccum;
    return i;
}

/* kissfft.hh:83-116 (kf_work), :128-135 (kf_bfly2), :137-157 (kf_bfly4).
 * Digit-reversed load (plan->rev), then the stages deepest-first. */
void lo_fft(const lo_plan *p, const lo_cf *in, lo_cf *F)
{
    const int N = p->n, L = p->nstages;
    const lo_cf *tw = p->twiddles;
    for (int j = 0; j < N; ++j) F[j] = in[p->rev[j]];
    for (int s = L - 1; s >= 0; --s) {
        const int r = p->radix[s], m = p->remainder[s];
        const int fstride = N / (r * m);
        for (int base = 0; base < N; base += r * m) {
            lo_cf *Fo = F + base;
            if (r == 2) {
                for (int k = 0; k < m; ++k) {
                    lo_cf t = cmul(Fo[m + k], tw[k * fstride]);
                    Fo[m + k] = csub(Fo[k], t);
                    Fo[k] = cadd(Fo[k], t);
                }
            } else {
                for (int k = 0; k < m; ++k) {
                    lo_cf s0 = cmul(Fo[k + m], tw[k * fstride]);
                    lo_cf s1 = cmul(Fo[k + 2 * m], tw[k * fstride * 2]);
                    lo_cf s2 = cmul(Fo[k + 3 * m], tw[k * fstride * 3]);
                    lo_cf s5 = csub(Fo[k], s1);
                    Fo[k] = cadd(Fo[k], s1);
                    lo_cf s3 = cadd(s0, s2);
                    lo_cf s4 = csub(s0, s2);
                    lo_cf s4r = { s4.im, -s4.re }; /* forward: negative_if_inverse=+1 */
                    Fo[k + 2 * m] = csub(Fo[k], s3);
                    Fo[k] = cadd(Fo[k], s3);
                    Fo[k + m] = cadd(s5, s4r);
                    Fo[k + 3 * m] = csub(s5, s4r);
                }
            }
        }
    }
}

/* LoRaDetector.hpp:29-64. */
size_t lo_detect(const lo_plan *p, const lo_cf *in, float *power,
                 float *power_avg, float *f_index, lo_cf *fft_out)
{
    const size_t N = (size_t)p->n;
    lo_cf *tmp = NULL;
    if (!fft_out) { tmp = (lo_cf *)malloc(sizeof(lo_cf) * N); fft_out = tmp; }
    lo_fft(p, in, fft_out);
    size_t maxIndex = 0;
    float maxValue = 0;
    double total = 0;
    for (size_t i = 0; i < N; i++) {
        float re = fft_out[i].re, im = fft_out[i].im;
        float mag2 = re * re + im * im;
        total += mag2;
        if (mag2 > maxValue) { maxIndex = i; maxValue = mag2; }
    }
    const float noise = sqrtf((float)(total - maxValue));
    const float fundamental = sqrtf(maxValue);
    *power_avg = 20 * log10f(noise) - p->power_scale;
    *power = 20 * log10f(fundamental) - p->power_scale;
    lo_cf l = fft_out[maxIndex > 0 ? maxIndex - 1 : N - 1];
    lo_cf r = fft_out[maxIndex < N - 1 ? maxIndex + 1 : 0];
    float left = cabsf(CMPLXF(l.re, l.im));   /* std::abs(complex<float>) */
    float right = cabsf(CMPLXF(r.re, r.im));
    const double demon = (2.0 * fundamental) - right - left;
    if (demon == 0.0) *f_index = 0.0;
    else *f_index = 0.5 * (right - left) / demon;
    free(tmp);
    return maxIndex;
}

/* LoRaDemod.cpp:160-162:  int -= float  =>  int(float(int) - float), one wrap. */
int32_t lo_fine_next(int32_t idx, float step, int32_t len)
{
    idx = (int32_t)((float)idx - step);
    if (idx < 0) idx += len;
    else if (idx >= len) idx -= len;
    return idx;
}

typedef struct {
    const lo_plan *p;
    const lo_cf *iq;
    size_t w0, w1, stride;
    const lo_cf *chirp, *fine;
    const int32_t *idx0;   /* per-window start index (pre-rolled) or NULL */
    float fine_step;
    lo_result *out;
    lo_cf *fft_out, *dec_out;
} lo_job;

static void *lo_worker(void *arg)
{
    lo_job *j = (lo_job *)arg;
    const lo_plan *p = j->p;
    const size_t N = (size_t)p->n;
    const int32_t flen = (int32_t)(N * 128);
    lo_cf *det = (lo_cf *)malloc(sizeof(lo_cf) * N);
    lo_cf *fft = (lo_cf *)malloc(sizeof(lo_cf) * N);
    for (size_t w = j->w0; w < j->w1; ++w) {
        const lo_cf *in = j->iq + w * j->stride;
        if (j->fine) {
            /* LoRaDemod.cpp:157-166 */
            int32_t idx = j->idx0[w];
            for (size_t i = 0; i < N; i++) {
                det[i] = cmul(cmul(in[i], j->chirp[i]), j->fine[idx]);
                idx = lo_fine_next(idx, j->fine_step, flen);
            }
        } else {
            /* TestDetector.cpp:25 form */
            for (size_t i = 0; i < N; i++) det[i] = cmul(in[i], j->chirp[i]);
        }
        if (j->dec_out) memcpy(j->dec_out + w * N, det, sizeof(lo_cf) * N);
        lo_result r;
        lo_cf *fo = j->fft_out ? j->fft_out + w * N : fft;
        r.index = (int32_t)lo_detect(p, det, &r.power, &r.power_avg, &r.f_index, fo);
        j->out[w] = r;
    }
    free(det);
    free(fft);
    return NULL;
}

void lo_demod_windows(const lo_plan *p, const lo_cf *iq, size_t n_windows,
                      size_t stride, const lo_cf *chirp, const lo_cf *fine,
                      int32_t *fine_idx, float fine_step, lo_result *out,
                      lo_cf *fft_out, lo_cf *dec_out, int nthreads)
{
    if (n_windows == 0) return;
    const size_t N = (size_t)p->n;
    int32_t *idx0 = NULL;
    if (fine) {
        /* pre-roll the sequential recurrence so windows can be split */
        idx0 = (int32_t *)malloc(sizeof(int32_t) * n_windows);
        int32_t idx = fine_idx ? *fine_idx : 0;
        const int32_t flen = (int32_t)(N * 128);
        for (size_t w = 0; w < n_windows; ++w) {
            idx0[w] = idx;
            /* with step == 0 the recurrence is the identity for in-range idx
             * (float(idx) is exact below 2^24), so skip the n steps */
            if (fine_step != 0.0f)
                for (size_t i = 0; i < N; ++i) idx = lo_fine_next(idx, fine_step, flen);
        }
        if (fine_idx) *fine_idx = idx;
    }
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n_windows) nthreads = (int)n_windows;
    lo_job *jobs = (lo_job *)calloc((size_t)nthreads, sizeof(lo_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; ++t) {
        lo_job *j = &jobs[t];
        j->p = p; j->iq = iq; j->stride = stride; j->chirp = chirp; j->fine = fine;
        j->idx0 = idx0; j->fine_step = fine_step; j->out = out;
        j->fft_out = fft_out; j->dec_out = dec_out;
        j->w0 = n_windows * (size_t)t / (size_t)nthreads;
        j->w1 = n_windows * (size_t)(t + 1) / (size_t)nthreads;
        if (nthreads == 1) lo_worker(j);
        else pthread_create(&th[t], NULL, lo_worker, j);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
    free(idx0);
}
