/*
 * lora_oracle.h -- CPU restatement of the LoRa-SDR demod hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * or as the timed CPU baseline.  The product path (lora-sdr_b200/csrc) never
 * links or calls this code and fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against
 * the reference's own headers compiled in place from /root/reference
 * (oracle/ref_shim.cpp -> oracle/_ref/libloraref.so, see tests/test_oracle_vs_ref.py)
 * and against the reference's only known-answer test on this path
 * (TestDetector.cpp:9-35), restated in tests/test_oracle_kat.py.  Fixtures
 * recorded from the real reference are committed under tests/golden/.
 *
 * Each function cites the reference file:line (relative to /root/reference)
 * whose arithmetic it restates.  Build with `-O2 -ffp-contract=off` and no
 * -march / -ffast-math (the reference's CMakeLists.txt:1-47 sets no flags).
 */
#ifndef LORA_ORACLE_H
#define LORA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } lo_cf;

/* One detector result; mirrors the outputs of LoRaDetector::detect
 * (LoRaDetector.hpp:29-64). */
typedef struct {
    int32_t index;     /* maxIndex                      LoRaDetector.hpp:33,45 */
    float   power;     /* 20log10(sqrt(max)) - scale    LoRaDetector.hpp:54    */
    float   power_avg; /* 20log10(sqrt(total-max))-scale LoRaDetector.hpp:53   */
    float   f_index;   /* fractional index              LoRaDetector.hpp:59-61 */
} lo_result;

#define LO_MAX_STAGES 16

typedef struct {
    int    sf;
    int    n;                       /* 2^sf */
    int    nstages;
    int    radix[LO_MAX_STAGES];    /* kissfft.hh:34-51: 4s, then 2s */
    int    remainder[LO_MAX_STAGES];/* m_s = n / (r_0 ... r_s)        */
    lo_cf *twiddles;                /* kissfft.hh:17-22, n entries    */
    float  power_scale;             /* LoRaDetector.hpp:18            */
    int   *rev;                     /* output slot j loads in[rev[j]] */
} lo_plan;

lo_plan *lo_plan_create(int sf);
void     lo_plan_destroy(lo_plan *p);

/* kissfft.hh:17-22  twiddle table, forward transform. */
void lo_twiddles(int n, lo_cf *tw);

/* LoRaDemod.cpp:97-107  conjugate-up / down dechirp tables (n entries each). */
void lo_tables_demod(int sf, lo_cf *up, lo_cf *down);

/* LoRaDemod.cpp:108-114  fine-tune table (n*128 entries). */
void lo_table_fine(int sf, lo_cf *fine);

/* ChirpGenerator.hpp:21-47  genChirp<float>. Returns samples generated. */
int lo_gen_chirp(lo_cf *samps, int N, int ovs, int NN, float f0, int down,
                 float ampl, float *phaseAccum);

/* kissfft.hh:77-157  forward FFT, natural-order output, unscaled. */
void lo_fft(const lo_plan *p, const lo_cf *in, lo_cf *out);

/* LoRaDetector.hpp:29-64  FFT + peak scan + dB metrics. `in` is the detector
 * input (what feed() stored); fft_out may be NULL. */
size_t lo_detect(const lo_plan *p, const lo_cf *in, float *power,
                 float *power_avg, float *f_index, lo_cf *fft_out);

/* One step of the fine-tune index recurrence, LoRaDemod.cpp:160-162.
 * step = _finefreqError * _fineSteps (float). */
int32_t lo_fine_next(int32_t idx, float step, int32_t table_len);

/*
 * Batched restatement of the dechirp loop + detect, LoRaDemod.cpp:157-172.
 *   iq        n_windows windows of n samples, window w starts at iq + w*stride
 *   chirp     dechirp table (n entries)
 *   fine      fine-tune table (n*128 entries) or NULL = single multiply
 *             (TestDetector.cpp:25 form: feed(i, chirp[i]*samp[i]))
 *   fine_idx  in: _fineTuneIndex at the first sample; out: value after the
 *             last processed sample (recurrence advances over the n processed
 *             samples of every window, in order). May be NULL when fine==NULL.
 *   fine_step _finefreqError*_fineSteps
 *   out       n_windows results
 *   fft_out / dec_out   optional n_windows*n dumps (fft port / dec port)
 *   nthreads  worker threads (static split of windows); with fine != NULL and
 *             fine_step != 0 the index recurrence is pre-rolled sequentially.
 */
void lo_demod_windows(const lo_plan *p, const lo_cf *iq, size_t n_windows,
                      size_t stride, const lo_cf *chirp, const lo_cf *fine,
                      int32_t *fine_idx, float fine_step, lo_result *out,
                      lo_cf *fft_out, lo_cf *dec_out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
