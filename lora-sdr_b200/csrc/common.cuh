// common.cuh -- device helpers shared by the demod kernels (sm_100a).
//
// Numerics contract ("exact mode"): the kernels reproduce the reference's float
// operand graph -- std::complex<float> multiply as four separately rounded
// products and two separately rounded sums (what GCC emits for finite operands),
// complex add/sub component-wise, |bin|^2 = re*re + im*im with three roundings.
// Every float op therefore goes through __fmul_rn/__fadd_rn/__fsub_rn, which
// nvcc never contracts into FMA, so FFT bins are bit-identical to kissfft's
// (kissfft.hh:128-157) regardless of where the values are stored.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/lorab200.h"

namespace lorab200 {

// Raw per-window detector outputs, finalised either on the device
// (finalize_device) or on the host with the host libm (bit-exact dB values).
struct __align__(16) RawResult {
    int32_t index;      // argmax bin
    float   max_mag2;   // |bin|^2 at the peak               LoRaDetector.hpp:46
    double  total;      // sum of |bin|^2 in double           LoRaDetector.hpp:42
    float2  left;       // bin[index-1 mod N]                 LoRaDetector.hpp:56
    float2  right;      // bin[index+1 mod N]                 LoRaDetector.hpp:57
};
static_assert(sizeof(RawResult) == 32, "RawResult must be 32 bytes");

struct DemodParams {
    const float2 *iq;         // windows: iq + w*stride
    unsigned long long n_windows;
    unsigned long long stride; // in samples
    const float2 *chirp;      // N entries (up or down table)
    const float2 *fine;       // N*128 entries or nullptr
    const int32_t *fine_x16;  // walking index: state at every 16th dechirped sample of the batch
                              // (fine_prefix.cuh), [n_windows * N/16] on the device; nullptr if step == 0
    int   fine_index0;        // the constant index when fine_step == 0
    float fine_step;          // _finefreqError * _fineSteps
    int   fine_len;           // N*128
    const float2 *tw;         // kissfft twiddles, N entries
    float power_scale;        // 20*log10(N) as float
    lorab200_result *out;     // device-finalised records, or nullptr
    RawResult *raw;           // raw records, or nullptr
    float2 *fft;              // optional [W][N]
    float2 *dec;              // optional [W][N]
    unsigned int early_iq;    // LORAB200_FLAG_INPUT_STABLE: IQ may be prefetched before the prior grid drains
    // FINE_WALK fast path (fine_walk.h): inside the region where every step of the recurrence is the
    // same integer fine_lin_step, a thread's 16 indices are x - i*step; valid while the indices stay in
    // [fine_lin_lo, fine_len).  fine_lin_ok == 0: always walk the recurrence literally.
    int fine_lin_ok, fine_lin_step, fine_lin_lo;
    unsigned int seg_windows; // multi-segment batches (lorab200_demod_multi_dev): windows per segment, else 0
};

// Multi-segment batches: the windows of a launch come from up to kMaxSegments independent
// device buffers (one per channel of a multi-channel receiver), seg_windows windows each;
// window w lives in segment w / seg_windows.  The base pointers travel as a kernel parameter
// (constant bank): no staging copy, safe under CUDA-graph capture.
constexpr int kMaxSegments = LORAB200_MAX_SEGMENTS;
template <bool MULTI> struct SegList { };
template <> struct SegList<true> { const float2 *p[kMaxSegments]; };


// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: every
// launcher keeps one flag per device (contexts on different devices, and on different host
// threads, are independent -- include/lorab200.h) and sets the attribute on first use there.
constexpr int kMaxDevices = 64;
struct SmemAttrOnce {
    std::atomic<bool> done[kMaxDevices];
    template <class Kernel>
    cudaError_t ensure(Kernel kernel, int bytes)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const bool tracked = dev >= 0 && dev < kMaxDevices;
        if (tracked && done[dev].load(std::memory_order_acquire)) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess && tracked) done[dev].store(true, std::memory_order_release);
        return e;
    }
};

// ---- exact complex arithmetic --------------------------------------------------
__device__ __forceinline__ float2 cx_mul(float2 a, float2 b)
{
    const float ac = __fmul_rn(a.x, b.x), bd = __fmul_rn(a.y, b.y);
    const float ad = __fmul_rn(a.x, b.y), bc = __fmul_rn(a.y, b.x);
    return make_float2(__fsub_rn(ac, bd), __fadd_rn(ad, bc));
}
__device__ __forceinline__ float2 cx_add(float2 a, float2 b)
{
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
__device__ __forceinline__ float2 cx_sub(float2 a, float2 b)
{
    return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y));
}
__device__ __forceinline__ float cx_mag2(float2 a)
{
    return __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y));
}

// kissfft.hh:137-157 kf_bfly4, forward transform, one butterfly on 4 points that
// have ALREADY been multiplied by their twiddles (a1..a3); a0 is untwiddled.
__device__ __forceinline__ void bfly4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    const float2 s5 = cx_sub(a0, a2);       // scratch[5] = F0 - s1   (s1 = twiddled F[2m])
    const float2 f0 = cx_add(a0, a2);       // F0 += s1
    const float2 s3 = cx_add(a1, a3);       // s0 + s2
    const float2 s4 = cx_sub(a1, a3);       // s0 - s2
    const float2 s4r = make_float2(s4.y, -s4.x);
    a2 = cx_sub(f0, s3);                    // F[2m] = F0 - s3
    a0 = cx_add(f0, s3);                    // F0 += s3
    a1 = cx_add(s5, s4r);                   // F[m]  = s5 + s4
    a3 = cx_sub(s5, s4r);                   // F[3m] = s5 - s4
}
// kissfft.hh:128-135 kf_bfly2 with the twiddle already applied to a1.
__device__ __forceinline__ void bfly2(float2 &a0, float2 &a1)
{
    const float2 t = a1;
    a1 = cx_sub(a0, t);
    a0 = cx_add(a0, t);
}

// One step of the fine-tune index recurrence (LoRaDemod.cpp:160-162):
// int -= float  ==>  int(float(int) - float), truncation, then a single wrap.
__device__ __forceinline__ int fine_next(int idx, float step, int len)
{
    idx = __float2int_rz(__fsub_rn(__int2float_rn(idx), step));
    if (idx < 0) idx += len;
    else if (idx >= len) idx -= len;
    return idx;
}

// (value, index) reduction rule that reproduces the sequential first-max scan
// of LoRaDetector.hpp:36-48 (strict '>' from a running max of 0): the larger
// value wins, equal values keep the smaller index.
__device__ __forceinline__ void argmax_merge(float &v, int &i, float ov, int oi)
{
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// glibc hypotf == (float)sqrt((double)x*x + (double)y*y) (verified on glibc 2.39);
// std::abs(std::complex<float>) calls it (LoRaDetector.hpp:56-57).
__device__ __forceinline__ float cx_abs(float2 a)
{
    const double x = double(a.x), y = double(a.y);
    return __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y))));
}

// LoRaDetector.hpp:50-61 on the device.  index and f_index are bit-exact.  power and
// power_avg: 20*log10(sqrt(x)) - scale is evaluated as 10*log10(2)*log2(x) - scale with one
// MUFU.LG2 per value (absolute error ~2^-22 in log2, i.e. about 1e-6 dB) where the reference
// calls glibc's sqrtf and log10f (not correctly rounded either): the device-resident path's
// contract is 1e-4 dB (tests/test_gpu_parity.py POWER_TOL_DB).  lorab200_demod_host finalises
// on the host with the host libm instead and is bit-exact.
__device__ __forceinline__ void finalize_device(const RawResult &r, float power_scale,
                                                lorab200_result *out, bool want_f_index = true)
{
    const float fundamental = __fsqrt_rn(r.max_mag2);
    lorab200_result o;
    o.index = r.index;
    const float kDbPerLog2 = 3.0102999566f; // 10*log10(2)
    o.power_avg = __fsub_rn(__fmul_rn(kDbPerLog2, __log2f(__double2float_rn(__dsub_rn(r.total, double(r.max_mag2))))), power_scale);
    o.power = __fsub_rn(__fmul_rn(kDbPerLog2, __log2f(r.max_mag2)), power_scale);
    // (the receiver kernels skip the fractional index -- two double square roots and a double division --
    //  for windows whose fIndex nothing reads: every state but FRAMESYNC, LoRaDemod.cpp:220)
    if (!want_f_index) { o.f_index = 0.0f; *out = o; return; }
    const float left = cx_abs(r.left), right = cx_abs(r.right);
    const double demon = __dsub_rn(__dsub_rn(__dmul_rn(2.0, double(fundamental)), double(right)), double(left));
    if (demon == 0.0) o.f_index = 0.0f;
    else o.f_index = __double2float_rn(__ddiv_rn(__dmul_rn(0.5, double(__fsub_rn(right, left))), demon));
    *out = o;
}

} // namespace lorab200
