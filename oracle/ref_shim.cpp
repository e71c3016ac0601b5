/*
 * ref_shim.cpp -- thin extern "C" wrapper around the REAL reference headers,
 * compiled in place from /root/reference (never copied into this repo) into
 * oracle/_ref/libloraref.so by oracle/Makefile.  TEST INFRASTRUCTURE ONLY.
 *
 * It exposes the same C entry points as lora_oracle.h with a `ref_` prefix so
 * the tests can compare the restatement bit-for-bit against the reference's
 * own code, and so bench.py can time the reference's own kissfft path as the
 * CPU baseline (cpu_baseline.kind == "reference").
 *
 * Uses:  LoRaDetector.hpp (LoRaDetector<float>::feed/detect), kissfft.hh,
 *        ChirpGenerator.hpp (genChirp<float>); <Pothos/Config.hpp> is an empty
 *        stub of ours (oracle/pothos_stub).  When built with -DREF_WITH_DEMOD
 *        it also compiles LoRaDemod.cpp (unmodified, against our ministream
 *        Pothos surface) to read the block's private dechirp / fine-tune tables.
 */
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <complex>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#ifdef REF_WITH_DEMOD
/* LoRaDemod.cpp includes LoRaDetector.hpp itself (the header has no include guard) */
#define private public
#include "LoRaDemod.cpp"      /* from /root/reference via -I, unmodified */
#undef private
#else
#include "LoRaDetector.hpp"   /* from /root/reference via -I */
#endif
#include "ChirpGenerator.hpp" /* from /root/reference via -I */

#include "lora_oracle.h"      /* lo_cf / lo_result layouts */

typedef std::complex<float> cf;

namespace {
// Persistent worker pool so the timed CPU baseline does not pay thread creation on
// every call (the reference itself is one thread per block; this is the fairest
// "all host cores" arrangement of its own per-window loop).
class Pool {
public:
    explicit Pool(int n) : n_(n), gen_(0), pending_(0), stop_(false)
    {
        for (int t = 0; t < n; ++t) th_.emplace_back([this, t] { run(t); });
    }
    ~Pool()
    {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return n_; }
    void parallel(const std::function<void(int)> &fn)
    {
        std::unique_lock<std::mutex> l(m_);
        fn_ = &fn; pending_ = n_; ++gen_;
        cv_.notify_all();
        done_.wait(l, [this] { return pending_ == 0; });
    }
private:
    void run(int t)
    {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int)> *fn;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(t);
            { std::lock_guard<std::mutex> l(m_); if (--pending_ == 0) done_.notify_all(); }
        }
    }
    int n_; unsigned long gen_; int pending_; bool stop_;
    const std::function<void(int)> *fn_ = nullptr;
    std::mutex m_; std::condition_variable cv_, done_;
    std::vector<std::thread> th_;
};
Pool *get_pool(int n)
{
    static Pool *pool = nullptr;
    if (!pool || pool->size() != n) { delete pool; pool = new Pool(n); }
    return pool;
}
} // namespace

extern "C" {

/* kissfft.hh:71-80 */
void ref_fft(int sf, const lo_cf *in, lo_cf *out)
{
    kissfft<float> fft(1 << sf, false);
    fft.transform(reinterpret_cast<const cf *>(in), reinterpret_cast<cf *>(out));
}

/* LoRaDetector.hpp:23-64 */
size_t ref_detect(int sf, const lo_cf *in, float *power, float *power_avg,
                  float *f_index, lo_cf *fft_out)
{
    const size_t N = size_t(1) << sf;
    LoRaDetector<float> det(N);
    for (size_t i = 0; i < N; i++) det.feed(i, cf(in[i].re, in[i].im));
    return det.detect(*power, *power_avg, *f_index, reinterpret_cast<cf *>(fft_out));
}

/* ChirpGenerator.hpp:21-47 */
int ref_gen_chirp(lo_cf *samps, int N, int ovs, int NN, float f0, int down,
                  float ampl, float *phaseAccum)
{
    return genChirp<float>(reinterpret_cast<cf *>(samps), N, ovs, NN, f0, down != 0, ampl, *phaseAccum);
}

/* The reference's per-window loop: N x feed(i, in[i]*chirp[i](*fine[idx])) then
 * detect()  (LoRaDemod.cpp:157-172 / TestDetector.cpp:23-26), run on
 * `nthreads` std::threads with one LoRaDetector<float> each.  The fine-tune
 * index recurrence (LoRaDemod.cpp:160-162) is applied verbatim with the same
 * operand types (int, float, size_t). */
void ref_demod_windows(int sf, const lo_cf *iq_, size_t n_windows, size_t stride,
                       const lo_cf *chirp_, const lo_cf *fine_, int32_t *fine_idx,
                       float finefreqError, lo_result *out, lo_cf *fft_out,
                       lo_cf *dec_out, int nthreads)
{
    const size_t N = size_t(1) << sf;
    const size_t _fineSteps = 128;
    const cf *iq = reinterpret_cast<const cf *>(iq_);
    const cf *chirp = reinterpret_cast<const cf *>(chirp_);
    const cf *fine = reinterpret_cast<const cf *>(fine_);
    if (nthreads < 1) nthreads = 1;

    /* one step of the index recurrence with the reference's operand types
     * (int -= float * size_t, then a single wrap; LoRaDemod.cpp:160-162) */
    auto advance = [N, _fineSteps](int &index, const float err) {
        index -= err * _fineSteps;
        const int len = int(N * _fineSteps);
        if (index < 0) index += len;
        else if (index >= len) index -= len;
    };

    std::vector<int> idx0(n_windows, 0);
    if (fine) {
        int _fineTuneIndex = fine_idx ? *fine_idx : 0;
        const float _finefreqError = finefreqError;
        for (size_t w = 0; w < n_windows; ++w) {
            idx0[w] = _fineTuneIndex;
            for (size_t i = 0; i < N; i++) advance(_fineTuneIndex, _finefreqError);
        }
        if (fine_idx) *fine_idx = _fineTuneIndex;
    }

    auto worker = [&](size_t w0, size_t w1) {
        LoRaDetector<float> det(N);
        std::vector<cf> fftbuf(N);
        for (size_t w = w0; w < w1; ++w) {
            const cf *in = iq + w * stride;
            if (fine) {
                int _fineTuneIndex = idx0[w];
                const float _finefreqError = finefreqError;
                for (size_t i = 0; i < N; i++) {
                    auto samp = in[i];
                    auto decd = samp * chirp[i] * fine[_fineTuneIndex];
                    advance(_fineTuneIndex, _finefreqError);
                    if (dec_out) std::memcpy(&dec_out[w * N + i], &decd, sizeof(cf));
                    det.feed(i, decd);
                }
            } else {
                for (size_t i = 0; i < N; i++) {
                    auto decd = in[i] * chirp[i];
                    if (dec_out) std::memcpy(&dec_out[w * N + i], &decd, sizeof(cf));
                    det.feed(i, decd);
                }
            }
            float power = 0, powerAvg = 0, fIndex = 0;
            cf *fo = fft_out ? reinterpret_cast<cf *>(fft_out) + w * N : fftbuf.data();
            size_t v = det.detect(power, powerAvg, fIndex, fo);
            out[w].index = int32_t(v);
            out[w].power = power;
            out[w].power_avg = powerAvg;
            out[w].f_index = fIndex;
        }
    };
    if (nthreads == 1 || n_windows < 2) { worker(0, n_windows); return; }
    Pool *pool = get_pool(nthreads);
    pool->parallel([&](int t) {
        worker(n_windows * size_t(t) / nthreads, n_windows * size_t(t + 1) / nthreads);
    });
}

#ifdef REF_WITH_DEMOD
/* Tables as built by the real LoRaDemod constructor, LoRaDemod.cpp:97-114. */
int ref_demod_tables(int sf, lo_cf *up, lo_cf *down, lo_cf *fine)
{
    LoRaDemod demod(sf);
    const size_t N = size_t(1) << sf;
    if (up) std::memcpy(up, demod._upChirpTable.data(), N * sizeof(cf));
    if (down) std::memcpy(down, demod._downChirpTable.data(), N * sizeof(cf));
    if (fine) std::memcpy(fine, demod._fineTuneTable.data(), N * 128 * sizeof(cf));
    return 1;
}
#else
int ref_demod_tables(int, lo_cf *, lo_cf *, lo_cf *) { return 0; }
#endif

} /* extern "C" */
