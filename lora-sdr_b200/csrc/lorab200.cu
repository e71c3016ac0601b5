// lorab200.cu -- context, C ABI (include/lorab200.h) and kernel dispatch.
//
// Host side of the drop-in boundary: everything a Pothos::Block::work() needs to
// replace LoRaDemod.cpp:157-172 with one call.  No torch types, no exceptions
// across the ABI, no CPU fallback: if CUDA is unavailable create() fails.
#include <cuda_runtime.h>
#include <sched.h>

#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <limits>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <new>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lorab200.h"
#include "common.cuh"
#include "demod_simple.cuh"
#ifndef LORAB200_NO_FUSED
#include "demod_fused.cuh"
#endif
#include "fine_walk.h"
#include "fine_prefix.cuh"
#include "mod_synth.cuh"
#include "decode.cuh"
#include "rx_kernel.cuh"
#include "rx_batch.cuh"
#include "acquire.cuh"
#include "finalize.h"
#include "tables.h"

using namespace lorab200;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

} // namespace

namespace {
// A few persistent helper threads per context for the host-side finalisation of large
// batches (libm log10f/hypotf per window); created on first use, joined in destroy.
class Helpers {
public:
    ~Helpers() { stop(); }
    void ensure(size_t n)
    {
        // a thread added later starts from the CURRENT generation: it must not take the pass that
        // was published before it existed (fn_ would be null / the counters stale)
        while (th_.size() < n) {
            const unsigned long g = gen_.load(std::memory_order_acquire);
            th_.emplace_back([this, g] { loop(g); });
        }
    }
    // run fn(i) for i in [0, n) on the helpers plus the calling thread
    void parallel(size_t n, const std::function<void(size_t)> &fn)
    {
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = &fn; next_ = 0; total_ = n; busy_ = th_.size();
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= n) break;
            fn(i);
        }
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return busy_ == 0; });
        fn_ = nullptr;
    }
    void stop()
    {
        { std::lock_guard<std::mutex> l(m_); quit_ = true; gen_.fetch_add(1, std::memory_order_release); }
        cv_.notify_all();
        for (auto &t : th_) t.join();
        th_.clear();
    }
private:
    void loop(unsigned long seen)
    {
        for (;;) {
            // Passes come in bursts (one per 6 MiB of a pageable batch, ~150 us apart): a helper that
            // has just finished one spins for a short while before it goes back to sleep on the
            // condition variable, whose wake-up (~100 us) is as long as the task it would be given.
            const auto spin_until = std::chrono::steady_clock::now() + std::chrono::microseconds(400);
            bool fresh = false;
            for (unsigned i = 0;; ++i) {
                if (gen_.load(std::memory_order_acquire) != seen) { fresh = true; break; }
                if ((i & 63u) == 63u && std::chrono::steady_clock::now() > spin_until) break;
                // (no PAUSE instruction here: under a hypervisor with pause-loop exiting it made a pass
                //  3x SLOWER than not spinning at all; a plain polling loop costs one core for <= 400 us)
            }
            const std::function<void(size_t)> *fn;
            size_t total;
            {
                std::unique_lock<std::mutex> l(m_);
                if (!fresh) cv_.wait(l, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                seen = gen_.load(std::memory_order_acquire);
                if (quit_) return;
                fn = fn_; total = total_;
            }
            for (;;) {
                const size_t i = next_.fetch_add(1);
                if (i >= total) break;
                (*fn)(i);
            }
            { std::lock_guard<std::mutex> l(m_); if (--busy_ == 0) done_.notify_all(); }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)> *fn_ = nullptr;
    std::atomic<size_t> next_{0};
    size_t total_ = 0, busy_ = 0;
    std::atomic<unsigned long> gen_{0};
    bool quit_ = false;
};
} // namespace

constexpr unsigned long long kFineSmallEntries = 8192; // fine prefix: closed form per entry up to here
constexpr size_t kRingSlots = 32, kRingSlotBytes = size_t(1) << 20; // pageable-input staging ring (32 MiB pinned)
constexpr size_t kRingFillThreads = 16;                                // slots filled concurrently (helper threads)
constexpr size_t kSmallBatchWindows = 16;          // lorab200_demod_host: zero-copy path up to here
constexpr size_t kSmallBatchBytes = size_t(256) << 10;

struct lorab200_ctx {
    Helpers helpers;
    int device = 0;
    int sf = 0;
    size_t n = 0;
    uint32_t flags = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;          // second lane of the host-buffer pipeline
    cudaStream_t stream_h2d = nullptr;       // all host->device chunk copies of the pipeline, back to back
    cudaEvent_t h2d_done[64] = {};           // chunk k's samples are on the device
    cudaEvent_t chunk_done[64] = {};          // D2H-complete events of the in-flight chunks
    std::string err;
    std::string last_kernel;
    unsigned long long launches = 0;
    float pscale = 0.0f;

    std::vector<cfloat> h_table[4];
    float2 *d_table[4] = {nullptr, nullptr, nullptr, nullptr};
#ifndef LORAB200_NO_FUSED
    Pass0Tw tw0;             // uniform pass-0 twiddles of the fused kernel (kernel parameter)
#endif

    // scratch for the host-buffer entry point and the fine-tune pre-roll
    DevBuf d_iq, d_raw, d_fft, d_dec;
    void *h_raw = nullptr;   // pinned
    size_t h_raw_cap = 0;
    void *h_stage = nullptr; // pinned staging of small batches (read by the kernel over PCIe)
    size_t h_stage_cap = 0;
    void *h_rx = nullptr;    // pinned: lorab200_rx_host's state / packets block (read and written by the kernel)
    size_t h_rx_cap = 0;
    DevBuf d_rx_iq;          // lorab200_rx_host: the channels' IQ on the device
    DevBuf d_acq_res, d_acq_valid, d_acq_tiles;   // lorab200_preamble_search_dev scratch
    DevBuf d_hio[4];         // host-buffer twins of encode / decode / mod: inputs and outputs on the device
    // pageable callers (a framework's ordinary port buffers): large H2D copies are staged through
    // a ring of pinned slots so that the PCIe copy of one slot overlaps the memcpy into the next
    void *h_ring = nullptr;
    cudaEvent_t ring_free[kRingSlots] = {};
    unsigned ring_next = 0;
    // walking fine-tune index (fine_prefix.cuh): T = g^16 for the cached step, two ping-pong
    // composition tables and the per-chunk states of the current batch
    DevBuf d_fine_t, d_fine_p[2], d_fine_x;
    DevBuf d_dec_scratch;                     // decoder: per-packet symbol / codeword / byte scratch
    DevBuf d_mod_phase, d_mod_phasor;         // modulator: REFERENCE-mode phases, IDEAL-mode exp(j*pi*m/N) table
    bool fine_t_valid = false;
    float fine_t_step = 0.0f;
    cudaEvent_t fine_done = nullptr;          // last kernel that read the fine scratch
    cudaEvent_t aux_done = nullptr;           // last kernel that used the modulator / codec scratch
};

namespace {

int fail(lorab200_ctx *c, int status, const char *what, cudaError_t e = cudaSuccess)
{
    std::string m = what;
    if (e != cudaSuccess) { m += ": "; m += cudaGetErrorString(e); }
    if (c) c->err = m; else g_create_error = m;
    return status;
}

#define CU_TRY(c, expr)                                                         \
    do {                                                                        \
        cudaError_t e__ = (expr);                                               \
        if (e__ != cudaSuccess) return fail((c), LORAB200_ERR_CUDA, #expr, e__); \
    } while (0)

int ensure_dev(lorab200_ctx *c, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return LORAB200_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(c, LORAB200_ERR_NOMEM, "cudaMalloc(scratch)", e); }
    b.cap = want;
    return LORAB200_OK;
}

int ensure_pinned(lorab200_ctx *c, void *&p, size_t &cap, size_t bytes)
{
    if (bytes <= cap) return LORAB200_OK;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(c, LORAB200_ERR_NOMEM, "cudaMallocHost(scratch)", e); }
    cap = want;
    return LORAB200_OK;
}

// CPUs this process may run on: respects the affinity mask a launcher gives each rank
// (bench.py binds every rank to its GPU's NUMA node), unlike hardware_concurrency()
size_t host_threads()
{
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return size_t(n);
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? h : 1;
}

// Slot fill: the destination (a pinned slot) is read next by the PCIe DMA engine, not by a CPU, so the
// copy writes it with non-temporal stores -- no read-for-ownership of the destination lines and no
// cache pollution (one thread: 16.5 vs 12.6 GB/s for memcpy on the build host).  Falls back to memcpy
// where AVX2 is absent or the slot is not 32-byte aligned.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) void stream_copy_avx2(char *d, const char *s, size_t n)
{
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 96), e);
    }
    if (i < n) std::memcpy(d + i, s + i, n - i);
    _mm_sfence();
}
#endif
void slot_fill(char *d, const char *s, size_t n)
{
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && std::getenv("LORAB200_NO_STREAM_COPY") == nullptr;
    if (avx2 && (reinterpret_cast<uintptr_t>(d) & 31u) == 0) { stream_copy_avx2(d, s, n); return; }
#endif
    std::memcpy(d, s, n);
}

bool is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// Host -> device copy of `bytes` on stream `s`.  Page-locked sources are copied directly;
// pageable ones go through the context's ring of pinned slots (memcpy into a slot, async copy
// out of it), which keeps the PCIe engine busy while the CPU fills the next slot.
int h2d_copy(lorab200_ctx *c, void *dst, const void *src, size_t bytes, cudaStream_t s, bool pageable)
{
    if (!pageable) {
        CU_TRY(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
        return LORAB200_OK;
    }
    if (!c->h_ring) {
        cudaError_t e = cudaMallocHost(&c->h_ring, kRingSlots * kRingSlotBytes);
        if (e != cudaSuccess) { (void)cudaGetLastError(); c->h_ring = nullptr; return fail(c, LORAB200_ERR_NOMEM, "cudaMallocHost(ring)", e); }
    }
    // Groups of up to `par` slots: the helper threads (and this one) memcpy one slot each, then the
    // slots' async copies are queued; the next group is filled while those copies run.  One thread
    // moves ~12 GB/s (measured, round 1), PCIe gen5 takes ~50: a handful of threads keep it busy.
    size_t par = host_threads() / 2;
    if (par > kRingFillThreads) par = kRingFillThreads;
    static const long env_par = std::getenv("LORAB200_STAGE_THREADS") ? std::atol(std::getenv("LORAB200_STAGE_THREADS")) : 0;
    if (env_par > 0 && size_t(env_par) < par) par = size_t(env_par);
    if (par < 1) par = 1;
    if (par > 1) c->helpers.ensure(par - 1);
    static const bool trace = std::getenv("LORAB200_TRACE") != nullptr;
    using clk = std::chrono::steady_clock;
    double t_wait = 0, t_fill = 0, t_enq = 0;
    for (size_t off = 0; off < bytes;) {
        const auto ta = clk::now();
        size_t n_slots = (bytes - off + kRingSlotBytes - 1) / kRingSlotBytes;
        if (n_slots > par) n_slots = par;
        unsigned slot[kRingFillThreads];
        for (size_t i = 0; i < n_slots; ++i) {
            slot[i] = c->ring_next++ % kRingSlots;
            CU_TRY(c, cudaEventSynchronize(c->ring_free[slot[i]])); // the copy that last used this slot is done
        }
        const auto tb = clk::now();
        const char *from = static_cast<const char *>(src) + off;
        const size_t left = bytes - off;
        char *ring = static_cast<char *>(c->h_ring);
        const std::function<void(size_t)> fill = [&](size_t i) {
            const size_t o = i * kRingSlotBytes, len = left - o < kRingSlotBytes ? left - o : kRingSlotBytes;
            slot_fill(ring + size_t(slot[i]) * kRingSlotBytes, from + o, len);
        };
        if (n_slots > 1) c->helpers.parallel(n_slots, fill);
        else fill(0);
        const auto tc = clk::now();
        for (size_t i = 0; i < n_slots; ++i) {
            const size_t o = i * kRingSlotBytes, len = left - o < kRingSlotBytes ? left - o : kRingSlotBytes;
            CU_TRY(c, cudaMemcpyAsync(static_cast<char *>(dst) + off + o, ring + size_t(slot[i]) * kRingSlotBytes, len,
                                      cudaMemcpyHostToDevice, s));
            CU_TRY(c, cudaEventRecord(c->ring_free[slot[i]], s));
        }
        off += n_slots * kRingSlotBytes;
        if (trace) {
            const auto td = clk::now();
            t_wait += std::chrono::duration<double>(tb - ta).count();
            t_fill += std::chrono::duration<double>(tc - tb).count();
            t_enq += std::chrono::duration<double>(td - tc).count();
        }
    }
    if (trace)
        std::fprintf(stderr, "lorab200 trace: staged %.1f MiB with %zu thread(s): slot wait %.0f us, fill %.0f us (%.1f GB/s), enqueue %.0f us\n",
                     bytes / 1048576.0, par, t_wait * 1e6, t_fill * 1e6, bytes / (t_fill > 0 ? t_fill : 1) / 1e9, t_enq * 1e6);
    return LORAB200_OK;
}

int upload_table(lorab200_ctx *c, int which)
{
    const size_t bytes = c->h_table[which].size() * sizeof(cfloat);
    if (!c->d_table[which]) CU_TRY(c, cudaMalloc(&c->d_table[which], bytes));
    CU_TRY(c, cudaMemcpy(c->d_table[which], c->h_table[which].data(), bytes, cudaMemcpyHostToDevice));
#ifndef LORAB200_NO_FUSED
    if (which == LORAB200_TABLE_TWIDDLE)
        fill_pass0_tw(c->sf, reinterpret_cast<const float2 *>(c->h_table[which].data()), c->tw0);
#endif
    return LORAB200_OK;
}

template <int SF, bool MULTI = false>
cudaError_t launch_simple_sf(const DemodParams &p, cudaStream_t s, int sm_count,
                             const SegList<MULTI> &segs = SegList<MULTI>())
{
    constexpr int N = 1 << SF;
    const size_t smem = size_t(N) * sizeof(float2) + size_t(N) * sizeof(int);
    static SmemAttrOnce attr;
    if (cudaError_t e = attr.ensure(demod_simple_kernel<SF, MULTI>, int(smem)); e != cudaSuccess) return e;
    unsigned long long grid = p.n_windows;
    const unsigned long long cap = (unsigned long long)sm_count * 64ull;
    if (grid > cap) grid = cap; // grid-stride over windows
    demod_simple_kernel<SF, MULTI><<<unsigned(grid), kSimpleThreads, smem, s>>>(p, segs);
    return cudaGetLastError();
}

cudaError_t launch_simple(int sf, const DemodParams &p, cudaStream_t s, int sm_count)
{
    switch (sf) {
    case 7: return launch_simple_sf<7>(p, s, sm_count);
    case 8: return launch_simple_sf<8>(p, s, sm_count);
    case 9: return launch_simple_sf<9>(p, s, sm_count);
    case 10: return launch_simple_sf<10>(p, s, sm_count);
    case 11: return launch_simple_sf<11>(p, s, sm_count);
    case 12: return launch_simple_sf<12>(p, s, sm_count);
    default: return cudaErrorInvalidValue;
    }
}

static_assert(sizeof(lorab200_rx_state) == sizeof(RxState) && sizeof(lorab200_rx_packet) == sizeof(RxPacketInfo),
              "the C structs mirror the kernel's");

template <int SF>
cudaError_t launch_rx_sf(const RxParams &p, cudaStream_t s, int sm_count)
{
    constexpr int N = 1 << SF;
    const size_t smem = size_t(N) * sizeof(float2) + size_t(N) * sizeof(int);
    static SmemAttrOnce attr;
    if (cudaError_t e = attr.ensure(rx_kernel<SF>, int(smem)); e != cudaSuccess) return e;
    unsigned grid = p.n_channels;
    const unsigned cap = unsigned(sm_count) * 16u;
    if (grid > cap) grid = cap;
    rx_kernel<SF><<<grid, kSimpleThreads, smem, s>>>(p);
    return cudaGetLastError();
}

template <int SF>
cudaError_t launch_rx_batch_sf(const RxParams &p, const Pass0Tw &tw0, cudaStream_t s, int sm_count)
{
    using K = RxBatchCfg<SF>;
    static SmemAttrOnce attr;
    if (cudaError_t e = attr.ensure(rx_batch_kernel<SF>, K::SMEM_BYTES); e != cudaSuccess) return e;
    unsigned grid = p.n_channels;
    const unsigned cap = unsigned(sm_count) * 8u;   // grid-stride over the channels beyond that
    if (grid > cap) grid = cap;
    rx_batch_kernel<SF><<<grid, K::CTA, K::SMEM_BYTES, s>>>(p, tw0);
    return cudaGetLastError();
}

bool stream_capturing(cudaStream_t s)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return st != cudaStreamCaptureStatusNone;
}

cudaError_t launch_simple_multi(int sf, const DemodParams &p, const SegList<true> &segs, cudaStream_t s, int sm_count)
{
    switch (sf) {
    case 7: return launch_simple_sf<7, true>(p, s, sm_count, segs);
    case 8: return launch_simple_sf<8, true>(p, s, sm_count, segs);
    case 9: return launch_simple_sf<9, true>(p, s, sm_count, segs);
    case 10: return launch_simple_sf<10, true>(p, s, sm_count, segs);
    case 11: return launch_simple_sf<11, true>(p, s, sm_count, segs);
    case 12: return launch_simple_sf<12, true>(p, s, sm_count, segs);
    default: return cudaErrorInvalidValue;
    }
}

// Per-context scratch (modulator phases, codec buffers) may still be in use by the previous call
// on another stream: order this stream after it / mark the end of this use.  (Not while `s` is
// being captured into a graph: a capture orders its own nodes.)
cudaError_t aux_scratch_acquire(lorab200_ctx *c, cudaStream_t s)
{
    return stream_capturing(s) ? cudaSuccess : cudaStreamWaitEvent(s, c->aux_done, 0);
}
cudaError_t aux_scratch_release(lorab200_ctx *c, cudaStream_t s)
{
    return stream_capturing(s) ? cudaSuccess : cudaEventRecord(c->aux_done, s);
}

// Walking fine-tune index (LoRaDemod.cpp:160-162): state at every 16th dechirped sample of
// the batch, on the device, by pointer jumping (fine_prefix.cuh).  Enqueued on `s`.
int fine_prefix(lorab200_ctx *c, int32_t x0, float step, int32_t len, size_t n_windows, cudaStream_t s)
{
    const unsigned long long q_end = (unsigned long long)n_windows * (c->n / kFineChunk);
    int st;
    if ((st = ensure_dev(c, c->d_fine_x, q_end * sizeof(int32_t)))) return st;
    if (q_end <= kFineSmallEntries) {
        // small batch: every entry in closed form, one tiny launch, no tables
        const FineWalk fw(len, step);
        FineWalkDev dv;
        dv.len = len; dv.d = step; dv.integer = fw.integer ? 1 : 0; dv.D = fw.D; dv.Kp = fw.Kp;
        if (!stream_capturing(s)) CU_TRY(c, cudaStreamWaitEvent(s, c->fine_done, 0));
        fine_small_kernel<<<unsigned((q_end + 127) / 128), 128, 0, s>>>(static_cast<int32_t *>(c->d_fine_x.p), dv, x0, q_end);
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        return LORAB200_OK;
    }
    for (DevBuf *b : {&c->d_fine_t, &c->d_fine_p[0], &c->d_fine_p[1]})
        if ((st = ensure_dev(c, *b, size_t(len) * sizeof(int32_t)))) return st;
    // the scratch may still be in use by the previous call on another stream (not while `s` is
    // being captured into a graph: a capture orders its own nodes)
    if (!stream_capturing(s)) CU_TRY(c, cudaStreamWaitEvent(s, c->fine_done, 0));
    const int threads = 256;
    auto blocks = [&](unsigned long long n) {
        unsigned long long b = (n + threads - 1) / threads;
        const unsigned long long cap = (unsigned long long)c->sm_count * 16ull;
        return unsigned(b < 1 ? 1 : (b > cap ? cap : b));
    };
    int32_t *T = static_cast<int32_t *>(c->d_fine_t.p);
    const bool capturing = stream_capturing(s);
    if (capturing || !c->fine_t_valid || c->fine_t_step != step) {
        fine_base_kernel<<<blocks(len), threads, 0, s>>>(T, step, len);
        CU_TRY(c, cudaGetLastError());
        // a launch captured into a graph only runs at replay: do not let eager calls rely on it
        c->fine_t_valid = !capturing;
        c->fine_t_step = step;
        c->launches++;
    }
    const int32_t *P = T;
    int32_t *X = static_cast<int32_t *>(c->d_fine_x.p);
    int round = 0;
    for (unsigned long long b = 1;; b *= kFineRadix, ++round) {
        const int compose = (b * kFineRadix < q_end) ? 1 : 0; // the last round needs no further table
        int32_t *Pn = static_cast<int32_t *>(c->d_fine_p[round & 1].p);
        const unsigned long long work = (compose && (unsigned long long)len > b) ? (unsigned long long)len : b;
        fine_round_kernel<<<blocks(work), threads, 0, s>>>(P, Pn, X, b, q_end, len, x0, compose);
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        if (!compose) break;
        P = Pn;
    }
    return LORAB200_OK;
}

// Common launch path. `results`/`raw` are device pointers (exactly one non-null).
int run_demod(lorab200_ctx *c, const void *iq_dev, size_t n_windows, size_t stride, int table,
              lorab200_fine *fine, lorab200_result *out_dev, RawResult *raw_dev, void *fft_dev,
              void *dec_dev, cudaStream_t s)
{
    if (table != LORAB200_TABLE_UP && table != LORAB200_TABLE_DOWN)
        return fail(c, LORAB200_ERR_INVALID, "table must be LORAB200_TABLE_UP or LORAB200_TABLE_DOWN");
    if (n_windows == 0) return LORAB200_OK;
    if (!iq_dev) return fail(c, LORAB200_ERR_INVALID, "iq pointer is null");
    if (!out_dev && !raw_dev) return fail(c, LORAB200_ERR_INVALID, "output pointer is null");
    if ((reinterpret_cast<uintptr_t>(iq_dev) & 7u) != 0)
        return fail(c, LORAB200_ERR_INVALID, "iq pointer must be 8-byte aligned");
    if (n_windows > 1 && stride == 0) return fail(c, LORAB200_ERR_INVALID, "stride_samples must be > 0");

    DemodParams p{};
    p.iq = static_cast<const float2 *>(iq_dev);
    p.n_windows = n_windows;
    p.stride = stride;
    p.chirp = c->d_table[table];
    p.tw = c->d_table[LORAB200_TABLE_TWIDDLE];
    p.power_scale = c->pscale;
    p.out = out_dev;
    p.raw = raw_dev;
    p.fft = static_cast<float2 *>(fft_dev);
    p.dec = static_cast<float2 *>(dec_dev);
    p.fine_len = int(c->n * kFineSteps);
    p.early_iq = (c->flags & LORAB200_FLAG_INPUT_STABLE) ? 1u : 0u;

    if (fine) {
        if (fine->index < 0 || fine->index >= p.fine_len)
            return fail(c, LORAB200_ERR_INVALID, "fine->index out of [0, N*128)");
        if (!(fine->step > -float(p.fine_len) && fine->step < float(p.fine_len)))
            return fail(c, LORAB200_ERR_INVALID, "|fine->step| must be below N*128");
        p.fine = c->d_table[LORAB200_TABLE_FINE];
        p.fine_step = fine->step;
    }

    const bool walk = fine && fine->step != 0.0f;
    if (fine && !walk) p.fine_index0 = fine->index;
    bool use_fused = false;
#ifndef LORAB200_NO_FUSED
    use_fused = !(c->flags & LORAB200_FLAG_SIMPLE_ONLY) && fused_applicable(c->sf, p);
#endif
    FineOrbit orbit;
    bool have_orbit = false;
    if (walk) {
        const FineWalk fwalk(p.fine_len, fine->step);
        // The fused kernel evaluates the index in O(1) per sample from the orbit of the walk
        // (fine_orbit.h) whenever the step has one: nothing to precompute on the device.  Otherwise
        // (steps within 2^-5 of an integer, the simple kernel): per-chunk index states by pointer
        // jumping (fine_prefix.cuh).  The end-of-batch index for the caller comes from the closed
        // form on the host (fine_walk.h) either way, so this call stays asynchronous.
        have_orbit = use_fused && build_fine_orbit(fwalk, fine->index, orbit);
        if (!have_orbit) {
            int st = fine_prefix(c, fine->index, fine->step, p.fine_len, n_windows, s);
            if (st) return st;
            p.fine_x16 = static_cast<const int32_t *>(c->d_fine_x.p);
        }
        fine->index = fwalk.advance(fine->index, uint64_t(n_windows) * c->n);
        // arithmetic-progression fast path of the kernel: an integer step d, or (rounding never carries,
        // Kp == none) every state with k = x - D >= 1 steps by D + 1
        const bool kp_none = fwalk.Kp == std::numeric_limits<int64_t>::max();
        p.fine_lin_ok = (fwalk.integer || kp_none) ? 1 : 0;
        p.fine_lin_step = int(fwalk.integer ? fwalk.D : fwalk.D + 1);
        p.fine_lin_lo = fwalk.integer ? 0 : int(fwalk.D + 1 > 0 ? fwalk.D + 1 : 0);
    }

    cudaError_t e;
#ifndef LORAB200_NO_FUSED
    if (use_fused) {
        // fine-tune with a zero step: the index never moves (LoRaDemod.cpp:160 with
        // _finefreqError == 0), so the whole batch is multiplied by one phasor
        float2 phasor = make_float2(1.f, 0.f);
        if (fine && !walk) {
            const cfloat f = c->h_table[LORAB200_TABLE_FINE][size_t(fine->index)];
            phasor = make_float2(f.re, f.im);
        }
        e = launch_fused(c->sf, p, c->tw0, walk ? (have_orbit ? FINE_ORBIT : FINE_WALK) : (fine ? FINE_CONST : FINE_NONE),
                         phasor, s, c->sm_count, have_orbit ? &orbit : nullptr);
        c->last_kernel = std::string("fused_sf") + std::to_string(c->sf);
    } else
#endif
    {
        e = launch_simple(c->sf, p, s, c->sm_count);
        c->last_kernel = std::string("simple_sf") + std::to_string(c->sf);
    }
    if (walk && !have_orbit && e == cudaSuccess && !stream_capturing(s)) e = cudaEventRecord(c->fine_done, s);
    if (e != cudaSuccess) return fail(c, LORAB200_ERR_CUDA, "kernel launch", e);
    c->launches++;
    return LORAB200_OK;
}

} // namespace

extern "C" {

int lorab200_abi_version(void) { return LORAB200_ABI_VERSION; }

const char *lorab200_strerror(int status)
{
    switch (status) {
    case LORAB200_OK: return "ok";
    case LORAB200_ERR_INVALID: return "invalid argument";
    case LORAB200_ERR_CUDA: return "CUDA error";
    case LORAB200_ERR_NOMEM: return "out of memory";
    case LORAB200_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
    }
}

const char *lorab200_last_error(const lorab200_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int lorab200_create(lorab200_ctx **out, int device, int sf, uint32_t flags)
{
    if (!out) return fail(nullptr, LORAB200_ERR_INVALID, "ctx out-pointer is null");
    *out = nullptr;
    if (sf < 7 || sf > 12) return fail(nullptr, LORAB200_ERR_INVALID, "sf must be in 7..12");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, LORAB200_ERR_CUDA, "no CUDA device (this library has no CPU fallback)", e);
    }
    if (device < 0 || device >= ndev) return fail(nullptr, LORAB200_ERR_INVALID, "device index out of range");
    lorab200_ctx *c = new (std::nothrow) lorab200_ctx();
    if (!c) return fail(nullptr, LORAB200_ERR_NOMEM, "ctx allocation");
    c->device = device;
    c->sf = sf;
    c->n = size_t(1) << sf;
    c->flags = flags;
    c->pscale = power_scale(sf);
    int st = LORAB200_OK;
    do {
        if ((e = cudaSetDevice(device)) != cudaSuccess) { st = fail(nullptr, LORAB200_ERR_CUDA, "cudaSetDevice", e); break; }
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { st = fail(nullptr, LORAB200_ERR_CUDA, "cudaGetDeviceProperties", e); break; }
        if (prop.major != 10) { st = fail(nullptr, LORAB200_ERR_UNSUPPORTED, "device is not sm_100 (B200); kernels are built for sm_100a only"); break; }
        c->sm_count = prop.multiProcessorCount;
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { st = fail(nullptr, LORAB200_ERR_CUDA, "cudaStreamCreate", e); break; }
        if ((e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking)) != cudaSuccess) { st = fail(nullptr, LORAB200_ERR_CUDA, "cudaStreamCreate", e); break; }
        if ((e = cudaStreamCreateWithFlags(&c->stream_h2d, cudaStreamNonBlocking)) != cudaSuccess) { st = fail(nullptr, LORAB200_ERR_CUDA, "cudaStreamCreate", e); break; }
        for (auto &ev : c->chunk_done)
            if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) break;
        for (auto &ev : c->h2d_done)
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->fine_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->aux_done, cudaEventDisableTiming);
        for (auto &ev : c->ring_free)
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e != cudaSuccess) { st = fail(nullptr, LORAB200_ERR_CUDA, "cudaEventCreate", e); break; }
        build_dechirp_tables(sf, c->h_table[LORAB200_TABLE_UP], c->h_table[LORAB200_TABLE_DOWN]);
        build_fine_table(sf, c->h_table[LORAB200_TABLE_FINE]);
        build_twiddles(sf, c->h_table[LORAB200_TABLE_TWIDDLE]);
        for (int t = 0; t < 4 && st == LORAB200_OK; ++t) {
            st = upload_table(c, t);
            if (st) g_create_error = c->err;
        }
        if (st == LORAB200_OK) {
            // modulator, IDEAL mode: exp(j*pi*m/N), m < 2N -- the 2N-point constellation every ideal chirp
            // sample lies on.  Built here (not on first use) so that lorab200_mod_dev never allocates or
            // copies synchronously: it stays legal inside a CUDA-graph capture.
            std::vector<float2> e(2 * c->n);
            for (size_t m = 0; m < e.size(); ++m) {
                const double a = M_PI * double(m) / double(c->n);
                e[m] = make_float2(float(std::cos(a)), float(std::sin(a)));
            }
            st = ensure_dev(c, c->d_mod_phasor, e.size() * sizeof(float2));
            if (st == LORAB200_OK &&
                cudaMemcpy(c->d_mod_phasor.p, e.data(), e.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess)
                st = fail(c, LORAB200_ERR_CUDA, "upload of the modulator phasor table");
            if (st) g_create_error = c->err;
        }
    } while (0);
    if (st != LORAB200_OK) { lorab200_destroy(c); return st; }
    *out = c;
    return LORAB200_OK;
}

void lorab200_destroy(lorab200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    if (c->stream_h2d) { cudaStreamSynchronize(c->stream_h2d); cudaStreamDestroy(c->stream_h2d); }
    for (auto &ev : c->h2d_done) if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->chunk_done) if (ev) cudaEventDestroy(ev);
    for (auto &t : c->d_table) if (t) cudaFree(t);
    for (DevBuf *b : {&c->d_iq, &c->d_raw, &c->d_fft, &c->d_dec, &c->d_fine_t, &c->d_fine_p[0], &c->d_fine_p[1], &c->d_fine_x, &c->d_mod_phase, &c->d_mod_phasor, &c->d_dec_scratch}) if (b->p) cudaFree(b->p);
    if (c->fine_done) cudaEventDestroy(c->fine_done);
    if (c->aux_done) cudaEventDestroy(c->aux_done);
    if (c->h_raw) cudaFreeHost(c->h_raw);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_rx) cudaFreeHost(c->h_rx);
    if (c->d_rx_iq.p) cudaFree(c->d_rx_iq.p);
    for (DevBuf &b : c->d_hio) if (b.p) cudaFree(b.p);
    for (DevBuf *b : {&c->d_acq_res, &c->d_acq_valid, &c->d_acq_tiles}) if (b->p) cudaFree(b->p);
    if (c->h_ring) cudaFreeHost(c->h_ring);
    for (auto &ev : c->ring_free) if (ev) cudaEventDestroy(ev);
    (void)cudaGetLastError();
    delete c;
}

int lorab200_sf(const lorab200_ctx *c) { return c ? c->sf : LORAB200_ERR_INVALID; }
size_t lorab200_n(const lorab200_ctx *c) { return c ? c->n : 0; }
int lorab200_device(const lorab200_ctx *c) { return c ? c->device : LORAB200_ERR_INVALID; }
unsigned long long lorab200_kernel_launches(const lorab200_ctx *c) { return c ? c->launches : 0; }
const char *lorab200_last_kernel(const lorab200_ctx *c) { return c ? c->last_kernel.c_str() : ""; }

const float *lorab200_table_host(const lorab200_ctx *c, int which, size_t *n_complex)
{
    if (!c || which < 0 || which > 3) return nullptr;
    if (n_complex) *n_complex = c->h_table[which].size();
    return reinterpret_cast<const float *>(c->h_table[which].data());
}

void *lorab200_table_dev(const lorab200_ctx *c, int which, size_t *n_complex)
{
    if (!c || which < 0 || which > 3) return nullptr;
    if (n_complex) *n_complex = c->h_table[which].size();
    return c->d_table[which];
}

int lorab200_set_table(lorab200_ctx *c, int which, const float *host_data)
{
    if (!c) return LORAB200_ERR_INVALID;
    if (which != LORAB200_TABLE_UP && which != LORAB200_TABLE_DOWN)
        return fail(c, LORAB200_ERR_INVALID, "only the UP / DOWN dechirp tables can be replaced");
    if (!host_data) return fail(c, LORAB200_ERR_INVALID, "host_data is null");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    std::memcpy(c->h_table[which].data(), host_data, c->n * sizeof(cfloat));
    return upload_table(c, which);
}

int lorab200_tables_sync_from_dev(lorab200_ctx *c)
{
    if (!c) return LORAB200_ERR_INVALID;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    for (int t = 0; t < 4; ++t)
        CU_TRY(c, cudaMemcpy(c->h_table[t].data(), c->d_table[t], c->h_table[t].size() * sizeof(cfloat),
                             cudaMemcpyDeviceToHost));
#ifndef LORAB200_NO_FUSED
    fill_pass0_tw(c->sf, reinterpret_cast<const float2 *>(c->h_table[LORAB200_TABLE_TWIDDLE].data()), c->tw0);
#endif
    return LORAB200_OK;
}

size_t lorab200_table_len(int sf, int which)
{
    if (sf < 7 || sf > 12 || which < 0 || which > 3) return 0;
    const size_t n = size_t(1) << sf;
    return which == LORAB200_TABLE_FINE ? n * kFineSteps : n;
}

int lorab200_build_table(int sf, int which, float *out)
{
    if (!out || lorab200_table_len(sf, which) == 0) return LORAB200_ERR_INVALID;
    std::vector<cfloat> a, b;
    switch (which) {
    case LORAB200_TABLE_UP: build_dechirp_tables(sf, a, b); break;
    case LORAB200_TABLE_DOWN: build_dechirp_tables(sf, b, a); break;
    case LORAB200_TABLE_FINE: build_fine_table(sf, a); break;
    default: build_twiddles(sf, a); break;
    }
    std::memcpy(out, a.data(), a.size() * sizeof(cfloat));
    return LORAB200_OK;
}

int32_t lorab200_fine_advance(int sf, int32_t index, float step, uint64_t n_samples)
{
    if (sf < 7 || sf > 12) return LORAB200_ERR_INVALID;
    const int32_t len = int32_t((size_t(1) << sf) * kFineSteps);
    if (index < 0 || index >= len || !(step > -float(len) && step < float(len))) return LORAB200_ERR_INVALID;
    if (step == 0.0f) return index;
    return FineWalk(len, step).advance(index, n_samples);
}

int lorab200_gen_chirp(float *samps, int N, int ovs, int NN, float f0, int down, float ampl,
                       float *phase_accum)
{
    if (!samps || !phase_accum || N <= 0 || ovs <= 0 || NN < 0) return LORAB200_ERR_INVALID;
    return gen_chirp(reinterpret_cast<cfloat *>(samps), N, ovs, NN, f0, down != 0, ampl, *phase_accum);
}

int lorab200_demod_dev(lorab200_ctx *c, const void *iq_dev, size_t n_windows, size_t stride_samples,
                       int table, lorab200_fine *fine, lorab200_result *out_dev, void *fft_dev,
                       void *dec_dev, void *cuda_stream)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream); // NULL == CUDA default stream
    return run_demod(c, iq_dev, n_windows, stride_samples, table, fine, out_dev, nullptr, fft_dev, dec_dev, s);
}

int lorab200_demod_multi_dev(lorab200_ctx *c, const void *const *iq_dev_list, size_t n_segments,
                             size_t windows_per_segment, size_t stride_samples, int table,
                             lorab200_result *out_dev, void *cuda_stream)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_segments == 0 || windows_per_segment == 0) return LORAB200_OK;
    if (!iq_dev_list || !out_dev) return fail(c, LORAB200_ERR_INVALID, "iq_dev_list / out_dev is null");
    if (n_segments > size_t(kMaxSegments)) return fail(c, LORAB200_ERR_UNSUPPORTED, "more than LORAB200_MAX_SEGMENTS segments");
    if (table != LORAB200_TABLE_UP && table != LORAB200_TABLE_DOWN)
        return fail(c, LORAB200_ERR_INVALID, "table must be LORAB200_TABLE_UP or LORAB200_TABLE_DOWN");
    if (windows_per_segment > 1 && stride_samples == 0) return fail(c, LORAB200_ERR_INVALID, "stride_samples must be > 0");
    if (n_segments * windows_per_segment >= (size_t(1) << 31) || windows_per_segment >= (size_t(1) << 31))
        return fail(c, LORAB200_ERR_UNSUPPORTED, "too many windows in one call");
    CU_TRY(c, cudaSetDevice(c->device));
    SegList<true> segs{};
    bool aligned16 = true;
    for (size_t i = 0; i < n_segments; ++i) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(iq_dev_list[i]);
        if (a == 0) return fail(c, LORAB200_ERR_INVALID, "a segment pointer is null");
        if (a & 7u) return fail(c, LORAB200_ERR_INVALID, "segment pointers must be 8-byte aligned");
        if (a & 15u) aligned16 = false;
        segs.p[i] = static_cast<const float2 *>(iq_dev_list[i]);
    }
    DemodParams p{};
    p.iq = segs.p[0];
    p.n_windows = n_segments * windows_per_segment;
    p.stride = stride_samples;
    p.chirp = c->d_table[table];
    p.tw = c->d_table[LORAB200_TABLE_TWIDDLE];
    p.power_scale = c->pscale;
    p.out = out_dev;
    p.fine_len = int(c->n * kFineSteps);
    p.early_iq = (c->flags & LORAB200_FLAG_INPUT_STABLE) ? 1u : 0u;
    p.seg_windows = unsigned(windows_per_segment);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    cudaError_t e;
#ifndef LORAB200_NO_FUSED
    if (!(c->flags & LORAB200_FLAG_SIMPLE_ONLY) && aligned16 && fused_applicable(c->sf, p)) {
        e = launch_fused_multi(c->sf, p, c->tw0, segs, s, c->sm_count);
        c->last_kernel = std::string("fused_multi_sf") + std::to_string(c->sf);
    } else
#endif
    {
        e = launch_simple_multi(c->sf, p, segs, s, c->sm_count);
        c->last_kernel = std::string("simple_multi_sf") + std::to_string(c->sf);
    }
    if (e != cudaSuccess) return fail(c, LORAB200_ERR_CUDA, "kernel launch", e);
    c->launches++;
    return LORAB200_OK;
}

int lorab200_rx_dev(lorab200_ctx *c, const void *iq_dev, size_t n_channels, size_t channel_stride_samples,
                    size_t n_samples, unsigned sync, float thresh_db, size_t mtu, lorab200_rx_state *state_dev,
                    int16_t *partial_dev, int16_t *symbols_dev, lorab200_rx_packet *info_dev, int32_t *n_packets_dev,
                    uint64_t *consumed_dev, size_t max_packets, void *cuda_stream)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_channels == 0) return LORAB200_OK;
    if (!iq_dev || !state_dev || !partial_dev || !symbols_dev || !info_dev || !n_packets_dev || !consumed_dev)
        return fail(c, LORAB200_ERR_INVALID, "null pointer");
    if ((reinterpret_cast<uintptr_t>(iq_dev) & 7u) != 0) return fail(c, LORAB200_ERR_INVALID, "iq pointer must be 8-byte aligned");
    if (sync > 0xff) return fail(c, LORAB200_ERR_INVALID, "sync is one byte");
    if (mtu == 0 || mtu > 32767) return fail(c, LORAB200_ERR_INVALID, "mtu must be 1..32767");
    if (max_packets == 0 || max_packets > (size_t(1) << 20)) return fail(c, LORAB200_ERR_INVALID, "max_packets out of range");
    if (n_channels > (size_t(1) << 24)) return fail(c, LORAB200_ERR_UNSUPPORTED, "too many channels");
    if (n_channels > 1 && channel_stride_samples < n_samples) return fail(c, LORAB200_ERR_INVALID, "channel_stride_samples < n_samples");
    CU_TRY(c, cudaSetDevice(c->device));
    RxParams p{};
    p.iq = static_cast<const float2 *>(iq_dev);
    p.chan_stride = channel_stride_samples;
    p.n_samples = n_samples;
    p.n_channels = unsigned(n_channels);
    p.up = c->d_table[LORAB200_TABLE_UP];
    p.down = c->d_table[LORAB200_TABLE_DOWN];
    p.fine = c->d_table[LORAB200_TABLE_FINE];
    p.tw = c->d_table[LORAB200_TABLE_TWIDDLE];
    p.power_scale = c->pscale;
    p.fine_len = int(c->n * kFineSteps);
    p.cfg.n = int(c->n); p.cfg.sync = int(sync); p.cfg.thresh = thresh_db; p.cfg.mtu = int(mtu);
    p.state = reinterpret_cast<RxState *>(state_dev);
    p.partial = partial_dev;
    p.symbols = symbols_dev;
    p.info = reinterpret_cast<RxPacketInfo *>(info_dev);
    p.n_packets = n_packets_dev;
    p.consumed = reinterpret_cast<unsigned long long *>(consumed_dev);
    p.max_packets = unsigned(max_packets);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    cudaError_t e;
    const bool simple = (c->flags & LORAB200_FLAG_SIMPLE_ONLY) != 0;   // the one-window-at-a-time cross-check kernel
    switch (c->sf) {
    case 7: e = simple ? launch_rx_sf<7>(p, s, c->sm_count) : launch_rx_batch_sf<7>(p, c->tw0, s, c->sm_count); break;
    case 8: e = simple ? launch_rx_sf<8>(p, s, c->sm_count) : launch_rx_batch_sf<8>(p, c->tw0, s, c->sm_count); break;
    case 9: e = simple ? launch_rx_sf<9>(p, s, c->sm_count) : launch_rx_batch_sf<9>(p, c->tw0, s, c->sm_count); break;
    case 10: e = simple ? launch_rx_sf<10>(p, s, c->sm_count) : launch_rx_batch_sf<10>(p, c->tw0, s, c->sm_count); break;
    case 11: e = simple ? launch_rx_sf<11>(p, s, c->sm_count) : launch_rx_batch_sf<11>(p, c->tw0, s, c->sm_count); break;
    default: e = simple ? launch_rx_sf<12>(p, s, c->sm_count) : launch_rx_batch_sf<12>(p, c->tw0, s, c->sm_count); break;
    }
    if (e != cudaSuccess) return fail(c, LORAB200_ERR_CUDA, "rx kernel launch", e);
    c->launches++;
    c->last_kernel = std::string(simple ? "rx_simple_sf" : "rx_sf") + std::to_string(c->sf);
    return LORAB200_OK;
}

int lorab200_rx_host(lorab200_ctx *c, const float *const *iq_list, size_t n_channels, size_t n_samples,
                     unsigned sync, float thresh_db, size_t mtu, lorab200_rx_state *state, int16_t *partial,
                     int16_t *symbols, lorab200_rx_packet *info, int32_t *n_packets, uint64_t *consumed,
                     size_t max_packets)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_channels == 0) return LORAB200_OK;
    if (!iq_list || !state || !partial || !symbols || !info || !n_packets || !consumed)
        return fail(c, LORAB200_ERR_INVALID, "null pointer");
    if (mtu == 0 || mtu > 32767) return fail(c, LORAB200_ERR_INVALID, "mtu must be 1..32767");
    if (max_packets == 0 || max_packets > 4096) return fail(c, LORAB200_ERR_INVALID, "max_packets must be 1..4096");
    if (n_channels > 4096) return fail(c, LORAB200_ERR_UNSUPPORTED, "at most 4096 channels per call");
    for (size_t ch = 0; ch < n_channels; ++ch)
        if (!iq_list[ch]) return fail(c, LORAB200_ERR_INVALID, "null channel pointer");
    CU_TRY(c, cudaSetDevice(c->device));
    // One pinned block holds everything small that crosses the boundary; the kernel reads and writes
    // it in place (page-locked memory is device-addressable), so a call is: stage the IQ, one launch,
    // one synchronise -- no copy calls for state, packets or symbols.
    const size_t o_state = 0;
    const size_t o_cons = o_state + ((n_channels * sizeof(lorab200_rx_state) + 15) & ~size_t(15));
    const size_t o_npk = o_cons + n_channels * sizeof(uint64_t);
    const size_t o_info = (o_npk + n_channels * sizeof(int32_t) + 15) & ~size_t(15);
    const size_t o_part = o_info + n_channels * max_packets * sizeof(lorab200_rx_packet);
    const size_t o_sym = (o_part + n_channels * mtu * sizeof(int16_t) + 15) & ~size_t(15);
    const size_t total = o_sym + n_channels * max_packets * mtu * sizeof(int16_t);
    int st;
    if ((st = ensure_pinned(c, c->h_rx, c->h_rx_cap, total))) return st;
    char *hb = static_cast<char *>(c->h_rx);
    std::memcpy(hb + o_state, state, n_channels * sizeof(lorab200_rx_state));
    std::memcpy(hb + o_part, partial, n_channels * mtu * sizeof(int16_t));
    const size_t stride = (n_samples + 1) & ~size_t(1);                 // keep every channel 16-byte aligned
    if ((st = ensure_dev(c, c->d_rx_iq, n_channels * stride * sizeof(float2)))) return st;
    float2 *d_iq = static_cast<float2 *>(c->d_rx_iq.p);
    for (size_t ch = 0; ch < n_channels; ++ch)
        if ((st = h2d_copy(c, d_iq + ch * stride, iq_list[ch], n_samples * sizeof(float2), c->stream, is_pageable(iq_list[ch]))))
            { cudaStreamSynchronize(c->stream); return st; }
    st = lorab200_rx_dev(c, d_iq, n_channels, stride, n_samples, sync, thresh_db, mtu,
                         reinterpret_cast<lorab200_rx_state *>(hb + o_state), reinterpret_cast<int16_t *>(hb + o_part),
                         reinterpret_cast<int16_t *>(hb + o_sym), reinterpret_cast<lorab200_rx_packet *>(hb + o_info),
                         reinterpret_cast<int32_t *>(hb + o_npk), reinterpret_cast<uint64_t *>(hb + o_cons),
                         max_packets, c->stream);
    if (st) { cudaStreamSynchronize(c->stream); return st; }
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    std::memcpy(state, hb + o_state, n_channels * sizeof(lorab200_rx_state));
    std::memcpy(partial, hb + o_part, n_channels * mtu * sizeof(int16_t));
    std::memcpy(consumed, hb + o_cons, n_channels * sizeof(uint64_t));
    std::memcpy(n_packets, hb + o_npk, n_channels * sizeof(int32_t));
    std::memcpy(info, hb + o_info, n_channels * max_packets * sizeof(lorab200_rx_packet));
    // the "power" / "snr" signals with the host libm (LoRaDetector.hpp:50-54, LoRaDemod.cpp:173): the
    // reference's own bits, from the raw peak and total the kernel kept for the DOWNCHIRP1 window
    auto host_db = [&](float max_mag2, double total, float &power, float &snr) {
        RawResultHost rr{};
        rr.max_mag2 = max_mag2;
        rr.total = total;
        lorab200_result o;
        finalize_host(&rr, 1, c->pscale, &o);
        power = o.power;
        snr = o.power - o.power_avg;
    };
    for (size_t ch = 0; ch < n_channels; ++ch) {
        const size_t np = size_t(n_packets[ch] < 0 ? 0 : n_packets[ch]);
        for (size_t q = 0; q < np; ++q) {
            lorab200_rx_packet &pk = info[ch * max_packets + q];
            host_db(pk.max_mag2, pk.total, pk.power, pk.snr);
        }
        if (state[ch].phase == RX_QUARTERCHIRP || state[ch].phase == RX_DATASYMBOLS)
            host_db(state[ch].sig_max_mag2, state[ch].sig_total, state[ch].sig_power, state[ch].sig_snr);
    }
    // only the packets that were written (the rest of the slots is scratch)
    for (size_t ch = 0; ch < n_channels; ++ch) {
        const size_t np = size_t(n_packets[ch] < 0 ? 0 : n_packets[ch]);
        if (np) std::memcpy(symbols + ch * max_packets * mtu, hb + o_sym + ch * max_packets * mtu * sizeof(int16_t),
                            np * mtu * sizeof(int16_t));
    }
    return LORAB200_OK;
}

int lorab200_decode_dev(lorab200_ctx *c, const uint16_t *symbols_dev, size_t n_packets, size_t symbols_stride,
                        const int32_t *n_symbols_dev, const lorab200_decoder_cfg *cfg, uint8_t *bytes_dev,
                        size_t bytes_stride, int32_t *lengths_dev, void *cuda_stream)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_packets == 0) return LORAB200_OK;
    if (!symbols_dev || !cfg || !bytes_dev || !lengths_dev) return fail(c, LORAB200_ERR_INVALID, "null pointer");
    if (cfg->rdd < 0 || cfg->rdd > 4) return fail(c, LORAB200_ERR_INVALID, "rdd must be 0..4");
    const int ppm = cfg->ppm ? cfg->ppm : c->sf;
    if (ppm < 5 || ppm > c->sf) return fail(c, LORAB200_ERR_INVALID, "ppm must be 0 or 5..SF (LoRaDecoder.cpp:202-203)");
    if (cfg->data_length < 0 || cfg->data_length > 65535) return fail(c, LORAB200_ERR_INVALID, "data_length out of range");
    if (symbols_stride == 0 || symbols_stride > 4096) return fail(c, LORAB200_ERR_UNSUPPORTED, "symbols_stride must be 1..4096");
    if (bytes_stride == 0 || bytes_stride > (size_t(1) << 20)) return fail(c, LORAB200_ERR_INVALID, "bytes_stride out of range");
    CU_TRY(c, cudaSetDevice(c->device));
    DecodeParams p{};
    p.symbols = symbols_dev;
    p.n_packets = n_packets;
    p.sym_stride = unsigned(symbols_stride);
    p.n_syms = n_symbols_dev;
    p.cfg.sf = c->sf; p.cfg.ppm = cfg->ppm; p.cfg.rdd = cfg->rdd; p.cfg.explicit_hdr = cfg->explicit_header;
    p.cfg.hdr = cfg->hdr; p.cfg.data_length = cfg->data_length; p.cfg.crcc = cfg->crcc; p.cfg.error_check = cfg->error_check;
    p.out = bytes_dev;
    p.out_stride = unsigned(bytes_stride);
    p.lengths = lengths_dev;
    p.max_sym = codec_num_symbols(p.cfg, p.sym_stride);
    p.max_cw = codec_num_codewords(p.cfg, p.sym_stride);
    p.max_bytes = codec_scratch_bytes(p.max_cw);
    const size_t sym_bytes = (size_t(p.max_sym) * 2 * n_packets + 15) & ~size_t(15);
    const size_t cw_bytes = (size_t(p.max_cw) * n_packets + 15) & ~size_t(15);
    int st;
    if ((st = ensure_dev(c, c->d_dec_scratch, sym_bytes + cw_bytes + size_t(p.max_bytes) * n_packets))) return st;
    char *base = static_cast<char *>(c->d_dec_scratch.p);
    p.sym_scratch = reinterpret_cast<uint16_t *>(base);
    p.cw_scratch = reinterpret_cast<uint8_t *>(base + sym_bytes);
    p.byte_scratch = reinterpret_cast<uint8_t *>(base + sym_bytes + cw_bytes);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    CU_TRY(c, aux_scratch_acquire(c, s));
    decode_kernel<<<unsigned((n_packets + 63) / 64), 64, 0, s>>>(p);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, aux_scratch_release(c, s));
    c->launches++;
    c->last_kernel = "decode";
    return LORAB200_OK;
}

size_t lorab200_encode_symbols(int sf, const lorab200_encoder_cfg *cfg, size_t payload_length)
{
    if (!cfg || sf < 7 || sf > 12 || cfg->rdd < 0 || cfg->rdd > 4 || payload_length > 255) return 0;
    const int ppm = cfg->ppm ? cfg->ppm : sf;
    if (ppm < 5 || ppm > sf) return 0;
    EncoderCfg e{sf, cfg->ppm, cfg->rdd, cfg->explicit_header, cfg->crc, cfg->whitening};
    return codec_enc_symbols(e, unsigned(payload_length));
}

int lorab200_encode_dev(lorab200_ctx *c, const uint8_t *payloads_dev, size_t n_packets, size_t payload_stride,
                        const int32_t *lengths_dev, const lorab200_encoder_cfg *cfg, uint16_t *symbols_dev,
                        size_t symbols_stride, int32_t *n_symbols_dev, void *cuda_stream)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_packets == 0) return LORAB200_OK;
    if (!payloads_dev || !cfg || !symbols_dev || !n_symbols_dev) return fail(c, LORAB200_ERR_INVALID, "null pointer");
    if (cfg->rdd < 0 || cfg->rdd > 4) return fail(c, LORAB200_ERR_INVALID, "rdd must be 0..4");
    const int ppm = cfg->ppm ? cfg->ppm : c->sf;
    if (ppm < 5 || ppm > c->sf) return fail(c, LORAB200_ERR_INVALID, "ppm must be 0 or 5..SF (LoRaEncoder.cpp:176-177)");
    if (payload_stride == 0 || payload_stride > 255) return fail(c, LORAB200_ERR_INVALID, "payload_stride must be 1..255 (the header length field is one byte)");
    if (symbols_stride == 0 || symbols_stride > 65535) return fail(c, LORAB200_ERR_INVALID, "symbols_stride out of range");
    CU_TRY(c, cudaSetDevice(c->device));
    EncodeParams p{};
    p.payloads = payloads_dev;
    p.n_packets = n_packets;
    p.payload_stride = unsigned(payload_stride);
    p.lengths = lengths_dev;
    p.cfg = EncoderCfg{c->sf, cfg->ppm, cfg->rdd, cfg->explicit_header, cfg->crc, cfg->whitening};
    p.symbols = symbols_dev;
    p.sym_stride = unsigned(symbols_stride);
    p.n_symbols = n_symbols_dev;
    p.max_bytes = p.payload_stride + 2;
    p.max_cw = codec_enc_codewords(p.cfg, p.payload_stride);
    int st;
    const size_t bytes_total = (size_t(p.max_bytes) * n_packets + 15) & ~size_t(15);
    if ((st = ensure_dev(c, c->d_dec_scratch, bytes_total + size_t(p.max_cw) * n_packets))) return st;
    p.byte_scratch = static_cast<uint8_t *>(c->d_dec_scratch.p);
    p.cw_scratch = p.byte_scratch + bytes_total;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    CU_TRY(c, aux_scratch_acquire(c, s));
    encode_kernel<<<unsigned((n_packets + 63) / 64), 64, 0, s>>>(p);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, aux_scratch_release(c, s));
    c->launches++;
    c->last_kernel = "encode";
    return LORAB200_OK;
}

size_t lorab200_mod_samples(int sf, int ovs, size_t syms_per_packet, size_t padding)
{
    if (sf < 7 || sf > 12 || ovs < 1 || ovs > 256) return 0;
    // LoRaMod.cpp:222-229: the padding state always emits at least one zero symbol
    return size_t(mod_samples_per_packet(1u << sf, unsigned(ovs), unsigned(syms_per_packet), unsigned(padding ? padding : 1)));
}

int lorab200_mod_dev(lorab200_ctx *c, const uint16_t *symbols_dev, size_t n_packets, size_t syms_per_packet,
                     unsigned sync, size_t padding, float ampl, int ovs, int mode, void *iq_dev, void *cuda_stream)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_packets == 0) return LORAB200_OK;
    if (!iq_dev || (!symbols_dev && syms_per_packet)) return fail(c, LORAB200_ERR_INVALID, "symbols_dev / iq_dev is null");
    if (ovs < 1 || ovs > 256) return fail(c, LORAB200_ERR_INVALID, "ovs must be in 1..256 (LoRaMod::setOvs)");
    if (mode != LORAB200_MOD_REFERENCE && mode != LORAB200_MOD_IDEAL) return fail(c, LORAB200_ERR_INVALID, "unknown mode");
    if (mode == LORAB200_MOD_IDEAL && ovs != 1) return fail(c, LORAB200_ERR_UNSUPPORTED, "LORAB200_MOD_IDEAL needs ovs == 1");
    if (sync > 0xff) return fail(c, LORAB200_ERR_INVALID, "sync is one byte");
    if (syms_per_packet >= (size_t(1) << 24) || padding >= (size_t(1) << 24))
        return fail(c, LORAB200_ERR_UNSUPPORTED, "packet too long");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    ModParams p{};
    p.symbols = symbols_dev;
    p.n_packets = n_packets;
    p.n_syms = unsigned(syms_per_packet);
    p.padding = unsigned(padding ? padding : 1);   // LoRaMod.cpp:222-229: at least one zero symbol
    p.sync = sync;
    p.ovs = unsigned(ovs);
    p.n = unsigned(c->n);
    p.ampl = ampl;
    p.samples_per_packet = mod_samples_per_packet(p.n, p.ovs, p.n_syms, p.padding);
    p.out = static_cast<float2 *>(iq_dev);
    const unsigned long long total = p.n_packets * p.samples_per_packet;
    if (p.samples_per_packet >= (1ull << 32)) return fail(c, LORAB200_ERR_UNSUPPORTED, "packet too long");
    int st;
    auto blocks = [&](unsigned long long n) {
        unsigned long long b = (n + 255) / 256;
        const unsigned long long cap = (unsigned long long)c->sm_count * 16ull;
        return unsigned(b < 1 ? 1 : (b > cap ? cap : b));
    };
    if (mode == LORAB200_MOD_IDEAL) {
        // two samples per 128-bit store (mod_synth.cuh): packets start on 16-byte boundaries
        // (samples_per_packet is a multiple of 32), so only the base needs checking
        if ((reinterpret_cast<uintptr_t>(iq_dev) & 15u) != 0)
            return fail(c, LORAB200_ERR_INVALID, "LORAB200_MOD_IDEAL needs a 16-byte aligned iq_dev");
        p.phasor = static_cast<const float2 *>(c->d_mod_phasor.p);
        const unsigned bpp = unsigned((p.samples_per_packet + kModIdealSpan - 1) / kModIdealSpan);
        const unsigned long long nb = p.n_packets * bpp;
        const unsigned long long cap = (unsigned long long)c->sm_count * 32ull;
        mod_ideal_kernel<<<unsigned(nb < cap ? nb : cap), kModIdealSpan / 2, 0, s>>>(p, unsigned(c->sf), bpp);
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        c->last_kernel = "mod_ideal";
        return LORAB200_OK;
    }
    if ((st = ensure_dev(c, c->d_mod_phase, total * sizeof(float)))) return st;
    p.phase = static_cast<float *>(c->d_mod_phase.p);
    CU_TRY(c, aux_scratch_acquire(c, s));
    mod_phase_kernel<<<unsigned((n_packets + 31) / 32), 32, 0, s>>>(p);
    CU_TRY(c, cudaGetLastError());
    mod_polar_kernel<<<blocks(total), 256, 0, s>>>(p);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, aux_scratch_release(c, s));
    c->launches += 2;
    c->last_kernel = "mod_reference";
    return LORAB200_OK;
}

int lorab200_demod_host(lorab200_ctx *c, const float *iq, size_t n_windows, size_t stride_samples,
                        int table, lorab200_fine *fine, lorab200_result *out, float *fft, float *dec)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_windows == 0) return LORAB200_OK;
    if (!iq || !out) return fail(c, LORAB200_ERR_INVALID, "iq / out pointer is null");
    if (n_windows > 1 && stride_samples == 0) return fail(c, LORAB200_ERR_INVALID, "stride_samples must be > 0");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t n = c->n;
    const size_t span = (n_windows - 1) * stride_samples + n; // samples touched
    const size_t dump = n_windows * n * sizeof(float2);
    int st;
    if ((st = ensure_pinned(c, c->h_raw, c->h_raw_cap, n_windows * sizeof(RawResult)))) return st;

    // Small batches (a block's work() in the sequential phases of the receiver: one or two
    // windows per call): latency, not bandwidth.  The span is staged in page-locked memory and
    // the kernel reads it straight over PCIe and writes the raw records straight back (pinned
    // memory is device-addressable under unified addressing): one launch and one stream
    // synchronise instead of H2D copy + launch + D2H copy + event round trips.
    if (!fft && !dec && n_windows <= kSmallBatchWindows && span * sizeof(float2) <= kSmallBatchBytes) {
        if ((st = ensure_pinned(c, c->h_stage, c->h_stage_cap, span * sizeof(float2) + 16))) return st;
        std::memcpy(c->h_stage, iq, span * sizeof(float2));
        RawResultHost *h_raw = static_cast<RawResultHost *>(c->h_raw);
        st = run_demod(c, c->h_stage, n_windows, stride_samples, table, fine, nullptr,
                       reinterpret_cast<RawResult *>(h_raw), nullptr, nullptr, c->stream);
        if (st) { cudaStreamSynchronize(c->stream); return st; }
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        finalize_host(h_raw, n_windows, c->pscale, out);
        return LORAB200_OK;
    }

    if ((st = ensure_dev(c, c->d_iq, span * sizeof(float2)))) return st;
    if ((st = ensure_dev(c, c->d_raw, n_windows * sizeof(RawResult)))) return st;
    if (fft && (st = ensure_dev(c, c->d_fft, dump))) return st;
    if (dec && (st = ensure_dev(c, c->d_dec, dump))) return st;

    // Software pipeline over chunks of windows on two streams: the H2D copy of chunk
    // i+1 overlaps the kernel and the D2H copy of chunk i, and the host-side
    // finalisation (libm) of earlier chunks; at most 64 chunks.
    // chunk = 8..16 MiB of IQ (measured on B200 / PCIe gen5, SF7 64 MiB batches: 2 MiB chunks
    // 42 GB/s, 4 MiB 48, 8 MiB 49.4, 16 MiB 47.5 -- profiles/r01_e2e_chunk_sweep.txt)
    size_t cw = 2048;
    const size_t wbytes = n * sizeof(float2);
    if (cw * wbytes < (size_t(8) << 20)) cw = (size_t(8) << 20) / wbytes;
    if (cw * wbytes > (size_t(16) << 20)) cw = (size_t(16) << 20) / wbytes;
    if (cw == 0) cw = 1;
    if ((n_windows + cw - 1) / cw > 64) cw = (n_windows + 63) / 64;
    // a walking fine-tune index uses per-context scratch for its prefix (fine_prefix.cuh): one chunk
    if (fine && fine->step != 0.0f) cw = n_windows;
    // Chunk boundaries.  Full chunks of cw windows, then a tapered tail (cw/2, cw/4, cw/8, cw/8): when the last
    // copy has landed only the last chunk's kernel, D2H and host finalisation remain, and during that tail
    // the PCIe link idles until the caller's next batch -- a short last chunk keeps it short.
    size_t cb[65];
    size_t n_chunks = 0;
    {
        size_t w = 0;
        const bool taper = !(fine && fine->step != 0.0f) && cw >= 64 && n_windows >= 4 * cw && (n_windows + cw - 1) / cw <= 60;
        cb[0] = 0;
        while (w < n_windows) {
            size_t step = cw;
            const size_t left = n_windows - w;
            if (taper && left <= cw) {
                static const size_t div[4] = {2, 4, 8, 8};
                for (int q = 0; q < 4 && w < n_windows; ++q) {
                    size_t sub = (cw / div[q] + 7) / 8 * 8;   // (multiples of 8 windows: tiles never straddle chunks)
                    if (q == 3 || sub > n_windows - w) sub = n_windows - w;
                    w += sub;
                    cb[++n_chunks] = w;
                }
                break;
            }
            if (step > left) step = left;
            w += step;
            cb[++n_chunks] = w;
        }
    }
    const float2 *h_iq = reinterpret_cast<const float2 *>(iq);
    float2 *d_iq = static_cast<float2 *>(c->d_iq.p);
    RawResult *d_raw = static_cast<RawResult *>(c->d_raw.p);
    RawResultHost *h_raw = static_cast<RawResultHost *>(c->h_raw);
    const bool pageable = is_pageable(iq);
    size_t copied_to = 0; // samples [0, copied_to) are already on the device
    for (size_t k = 0; k < n_chunks; ++k) {
        cudaStream_t s = (k & 1) ? c->stream2 : c->stream;
        const size_t w0 = cb[k], w1 = cb[k + 1], nw = w1 - w0;
        const size_t s0 = w0 * stride_samples, s1 = (w1 - 1) * stride_samples + n; // sample range of the chunk
        const size_t from = s0 > copied_to ? s0 : copied_to;                        // skip what an earlier chunk brought
        // Every chunk's copy goes on ONE stream, so that the copy engine runs them back to back (with the copies
        // on the two compute streams each one started a few microseconds after the previous lane's dependency
        // resolved); the chunk's kernel waits for its copy -- and, the copies being in order, for everything an
        // earlier chunk brought (stride < N: the head of this chunk).
        if (s1 > from) {
            if ((st = h2d_copy(c, d_iq + from, h_iq + from, (s1 - from) * sizeof(float2), c->stream_h2d, pageable))) {
                cudaStreamSynchronize(c->stream_h2d); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->stream2);
                return st;
            }
            copied_to = s1;
        }
        CU_TRY(c, cudaEventRecord(c->h2d_done[k % 64], c->stream_h2d));
        CU_TRY(c, cudaStreamWaitEvent(s, c->h2d_done[k % 64], 0));
        st = run_demod(c, d_iq + s0, nw, stride_samples, table, fine, nullptr, d_raw + w0,
                       fft ? static_cast<float2 *>(c->d_fft.p) + w0 * n : nullptr,
                       dec ? static_cast<float2 *>(c->d_dec.p) + w0 * n : nullptr, s);
        if (st) { cudaStreamSynchronize(c->stream_h2d); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->stream2); return st; }
        CU_TRY(c, cudaMemcpyAsync(h_raw + w0, d_raw + w0, nw * sizeof(RawResult), cudaMemcpyDeviceToHost, s));
        if (fft) CU_TRY(c, cudaMemcpyAsync(reinterpret_cast<float2 *>(fft) + w0 * n, static_cast<float2 *>(c->d_fft.p) + w0 * n,
                                           nw * n * sizeof(float2), cudaMemcpyDeviceToHost, s));
        if (dec) CU_TRY(c, cudaMemcpyAsync(reinterpret_cast<float2 *>(dec) + w0 * n, static_cast<float2 *>(c->d_dec.p) + w0 * n,
                                           nw * n * sizeof(float2), cudaMemcpyDeviceToHost, s));
        CU_TRY(c, cudaEventRecord(c->chunk_done[k % 64], s));
    }
    // Host-side finalisation (libm log10f/hypotf, LoRaDetector.hpp:50-61) of chunk k as
    // soon as its results have landed, while later chunks are still in flight.  Large
    // batches split the chunks over a few threads so this stays hidden behind the H2D copy.
    std::atomic<int> ferr_code{int(cudaSuccess)};
    // finalisation tasks are slices of a chunk: <= 1024 windows, <= 128 in the last two chunks so that the
    // tail after the last copy is spread evenly over the helper threads
    struct FTask { uint32_t chunk; size_t p0, p1; };
    std::vector<FTask> tasks;
    for (size_t k = 0; k < n_chunks; ++k) {
        const size_t ft = (k + 2 >= n_chunks && n_chunks >= 4) ? 128 : 1024;
        for (size_t p0 = cb[k]; p0 < cb[k + 1]; p0 += ft)
            tasks.push_back({uint32_t(k), p0, p0 + ft < cb[k + 1] ? p0 + ft : cb[k + 1]});
    }
    const std::function<void(size_t)> finish = [&](size_t task) {
        const FTask &t = tasks[task];
        cudaError_t e = cudaEventSynchronize(c->chunk_done[t.chunk % 64]);
        if (e != cudaSuccess) { ferr_code = int(e); return; }
        finalize_host(h_raw + t.p0, t.p1 - t.p0, c->pscale, out + t.p0);
    };
    const size_t n_tasks = tasks.size();
    if (n_windows >= 16384 && n_chunks >= 4) {
        size_t n_thr = host_threads() / 2;   // (a rank bound to its share of a NUMA node has ~16 CPUs)
        if (n_thr > 16) n_thr = 16;          // (the tail after the last chunk is one pass of these threads)
        if (n_thr > n_tasks) n_thr = n_tasks;
        if (n_thr < 2) n_thr = 2;
        c->helpers.ensure(n_thr - 1);
        c->helpers.parallel(n_tasks, finish);   // tasks are taken in order, so in completion order
    } else {
        for (size_t k = 0; k < n_tasks; ++k) finish(k);
    }
    const cudaError_t ferr = cudaError_t(ferr_code.load());
    if (ferr != cudaSuccess) return fail(c, LORAB200_ERR_CUDA, "cudaEventSynchronize(chunk)", ferr);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream2));
    return LORAB200_OK;
}

int lorab200_preamble_search_dev(lorab200_ctx *c, const void *iq_dev, size_t n_samples, size_t stride_samples,
                                 float thresh_db, int min_windows, lorab200_preamble *found_dev,
                                 int32_t *n_found_dev, uint32_t max_found, void *cuda_stream)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (!n_found_dev) return fail(c, LORAB200_ERR_INVALID, "n_found_dev is null");
    if (max_found > 0 && !found_dev) return fail(c, LORAB200_ERR_INVALID, "found_dev is null");
    if (stride_samples == 0) return fail(c, LORAB200_ERR_INVALID, "stride_samples must be > 0");
    if (min_windows < 1) return fail(c, LORAB200_ERR_INVALID, "min_windows must be >= 1");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const size_t n = c->n;
    if (n_samples < n) {
        CU_TRY(c, cudaMemsetAsync(n_found_dev, 0, sizeof(int32_t), s));
        return LORAB200_OK;
    }
    if (!iq_dev) return fail(c, LORAB200_ERR_INVALID, "iq pointer is null");
    const size_t n_win = (n_samples - n) / stride_samples + 1;
    if (n_win >= (size_t(1) << 31)) return fail(c, LORAB200_ERR_UNSUPPORTED, "too many windows in one call");
    const size_t n_tiles = (n_win + kAcqTile - 1) / kAcqTile;
    int st;
    if ((st = ensure_dev(c, c->d_acq_res, n_win * sizeof(lorab200_result)))) return st;
    if ((st = ensure_dev(c, c->d_acq_valid, n_win))) return st;
    if ((st = ensure_dev(c, c->d_acq_tiles, 2 * n_tiles * sizeof(int)))) return st;
    lorab200_result *res = static_cast<lorab200_result *>(c->d_acq_res.p);
    // every candidate offset in one launch of the demod kernel (windows overlap: the batch is read from L2)
    if ((st = run_demod(c, iq_dev, n_win, stride_samples, LORAB200_TABLE_UP, nullptr, res, nullptr, nullptr, nullptr, s))) return st;
    AcqParams p{};
    p.res = res;
    p.n_win = n_win;
    p.stride = stride_samples;
    p.n = int(n);
    p.thresh = thresh_db;
    p.min_windows = min_windows;
    p.valid = static_cast<unsigned char *>(c->d_acq_valid.p);
    p.tile_count = static_cast<int *>(c->d_acq_tiles.p);
    p.tile_base = p.tile_count + n_tiles;
    p.n_tiles = unsigned(n_tiles);
    p.found = found_dev;
    p.n_found = n_found_dev;
    p.max_found = max_found;
    const unsigned grid = unsigned(n_tiles < size_t(c->sm_count) * 8 ? n_tiles : size_t(c->sm_count) * 8);
    acq_mark_kernel<<<grid, kAcqTile, 0, s>>>(p);
    acq_scan_kernel<<<1, 1024, 0, s>>>(p);
    acq_emit_kernel<<<grid, kAcqTile, 0, s>>>(p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, LORAB200_ERR_CUDA, "preamble search launch", e);
    c->launches += 3;
    return LORAB200_OK;
}

void *lorab200_alloc_pinned(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}

void lorab200_free_pinned(void *p)
{
    if (p) cudaFreeHost(p);
}

// ---- host-buffer twins of the packet-level entry points: what a block's work() calls ---------------
// (stage the batch, one launch, copy the results back, synchronise; the batch is a handful of packets)

int lorab200_encode_host(lorab200_ctx *c, const uint8_t *payloads, size_t n_packets, size_t payload_stride,
                         const int32_t *lengths, const lorab200_encoder_cfg *cfg, uint16_t *symbols,
                         size_t symbols_stride, int32_t *n_symbols)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_packets == 0) return LORAB200_OK;
    if (!payloads || !cfg || !symbols || !n_symbols) return fail(c, LORAB200_ERR_INVALID, "null pointer");
    CU_TRY(c, cudaSetDevice(c->device));
    int st;
    const size_t pb = n_packets * payload_stride, sb = n_packets * symbols_stride * sizeof(uint16_t), lb = n_packets * sizeof(int32_t);
    if ((st = ensure_dev(c, c->d_hio[0], pb)) || (st = ensure_dev(c, c->d_hio[1], lb)) ||
        (st = ensure_dev(c, c->d_hio[2], sb)) || (st = ensure_dev(c, c->d_hio[3], lb))) return st;
    CU_TRY(c, cudaMemcpyAsync(c->d_hio[0].p, payloads, pb, cudaMemcpyHostToDevice, c->stream));
    if (lengths) CU_TRY(c, cudaMemcpyAsync(c->d_hio[1].p, lengths, lb, cudaMemcpyHostToDevice, c->stream));
    st = lorab200_encode_dev(c, static_cast<const uint8_t *>(c->d_hio[0].p), n_packets, payload_stride,
                             lengths ? static_cast<const int32_t *>(c->d_hio[1].p) : nullptr, cfg,
                             static_cast<uint16_t *>(c->d_hio[2].p), symbols_stride, static_cast<int32_t *>(c->d_hio[3].p), c->stream);
    if (st) { cudaStreamSynchronize(c->stream); return st; }
    CU_TRY(c, cudaMemcpyAsync(symbols, c->d_hio[2].p, sb, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(n_symbols, c->d_hio[3].p, lb, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return LORAB200_OK;
}

int lorab200_decode_host(lorab200_ctx *c, const uint16_t *symbols, size_t n_packets, size_t symbols_stride,
                         const int32_t *n_symbols, const lorab200_decoder_cfg *cfg, uint8_t *bytes,
                         size_t bytes_stride, int32_t *lengths)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_packets == 0) return LORAB200_OK;
    if (!symbols || !cfg || !bytes || !lengths) return fail(c, LORAB200_ERR_INVALID, "null pointer");
    CU_TRY(c, cudaSetDevice(c->device));
    int st;
    const size_t sb = n_packets * symbols_stride * sizeof(uint16_t), bb = n_packets * bytes_stride, lb = n_packets * sizeof(int32_t);
    if ((st = ensure_dev(c, c->d_hio[0], sb)) || (st = ensure_dev(c, c->d_hio[1], lb)) ||
        (st = ensure_dev(c, c->d_hio[2], bb)) || (st = ensure_dev(c, c->d_hio[3], lb))) return st;
    CU_TRY(c, cudaMemcpyAsync(c->d_hio[0].p, symbols, sb, cudaMemcpyHostToDevice, c->stream));
    if (n_symbols) CU_TRY(c, cudaMemcpyAsync(c->d_hio[1].p, n_symbols, lb, cudaMemcpyHostToDevice, c->stream));
    st = lorab200_decode_dev(c, static_cast<const uint16_t *>(c->d_hio[0].p), n_packets, symbols_stride,
                             n_symbols ? static_cast<const int32_t *>(c->d_hio[1].p) : nullptr, cfg,
                             static_cast<uint8_t *>(c->d_hio[2].p), bytes_stride, static_cast<int32_t *>(c->d_hio[3].p), c->stream);
    if (st) { cudaStreamSynchronize(c->stream); return st; }
    CU_TRY(c, cudaMemcpyAsync(bytes, c->d_hio[2].p, bb, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(lengths, c->d_hio[3].p, lb, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return LORAB200_OK;
}

int lorab200_mod_host(lorab200_ctx *c, const uint16_t *symbols, size_t n_packets, size_t syms_per_packet, unsigned sync,
                      size_t padding, float ampl, int ovs, int mode, float *iq)
{
    if (!c) return LORAB200_ERR_INVALID;
    c->err.clear();
    if (n_packets == 0) return LORAB200_OK;
    if (!iq || (!symbols && syms_per_packet)) return fail(c, LORAB200_ERR_INVALID, "symbols / iq is null");
    if (ovs < 1 || ovs > 256) return fail(c, LORAB200_ERR_INVALID, "ovs must be in 1..256 (LoRaMod::setOvs)");
    CU_TRY(c, cudaSetDevice(c->device));
    int st;
    const size_t per = lorab200_mod_samples(c->sf, ovs, syms_per_packet, padding);
    const size_t sb = n_packets * (syms_per_packet ? syms_per_packet : 1) * sizeof(uint16_t), qb = n_packets * per * sizeof(float2);
    if ((st = ensure_dev(c, c->d_hio[0], sb)) || (st = ensure_dev(c, c->d_hio[2], qb))) return st;
    if (syms_per_packet) CU_TRY(c, cudaMemcpyAsync(c->d_hio[0].p, symbols, n_packets * syms_per_packet * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    st = lorab200_mod_dev(c, static_cast<const uint16_t *>(c->d_hio[0].p), n_packets, syms_per_packet, sync, padding, ampl, ovs,
                          mode, c->d_hio[2].p, c->stream);
    if (st) { cudaStreamSynchronize(c->stream); return st; }
    CU_TRY(c, cudaMemcpyAsync(iq, c->d_hio[2].p, qb, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return LORAB200_OK;
}

int lorab200_is_pinned(const void *p)
{
    return (p && !is_pageable(p)) ? 1 : 0;
}

void lorab200_pack_symbols(const lorab200_result *results, size_t n, int16_t *symbols)
{
    if (!results || !symbols) return;
    for (size_t i = 0; i < n; ++i) symbols[i] = int16_t(results[i].index);
}

} // extern "C"
