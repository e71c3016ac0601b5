// packed.cuh -- exact-mode complex arithmetic on Blackwell's packed FP32 pipe.
//
// sm_100 adds mul/add/fma.rn.f32x2 (SASS FMUL2 / FADD2 / FFMA2): one instruction
// performs two IEEE round-to-nearest FP32 operations on an aligned register pair,
// and each source operand takes a free modifier (swap halves, per-half sign,
// scalar broadcast).  A complex number kept as the pair (re, im) therefore costs
//   complex add / sub            1 instruction   (was 2)
//   multiply by -j               0 instructions  (operand modifier)
//   complex multiply             3 instructions  (was 6: 4 FMUL + 2 FADD)
// while every individual rounding is the same as in the scalar code, so results
// stay bit-identical to the reference's std::complex<float> arithmetic
// (tools/ubench/f32x2.cu measures the throughput: FMUL2+FADD2 interleaved
// sustain 256 lane-flops/clk/SM against 128 for scalar code).
#pragma once
#include <cuda_runtime.h>

namespace lorab200 {

typedef unsigned long long c64; // packed complex: low word = re, high word = im

__device__ __forceinline__ c64 cpack(float re, float im)
{
    c64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im));
    return r;
}
__device__ __forceinline__ float2 cunpack(c64 p)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ c64 p_mul(c64 a, c64 b)
{
    c64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c64 p_add(c64 a, c64 b)
{
    c64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c64 p_sub(c64 a, c64 b)
{
    c64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c64 p_fma(c64 a, c64 b, c64 c)
{
    c64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// (re, im) * (c, d) = (re*c - im*d, re*d + im*c) with each product and each sum
// rounded separately -- the operand graph of std::complex<float>::operator*.
//   P = (re*c, im*c)   Q = (im*d, re*d)   result = (P.x - Q.x, P.y + Q.y)
// The combine is written as an FMA by (-1, +1): fma(q, -1, p) rounds p - q once,
// exactly like the subtraction (q * -1 is exact).  It must NOT be written as
// add.rn.f32x2(P, +-Q): ptxas (12.9) contracts mul.rn.f32x2 feeding add.rn.f32x2
// into one FFMA2 even with explicit .rn and -fmad=false, which drops the rounding
// of the product.  With the FMA form there is nothing left to contract: P feeds an
// addend, Q a multiplicand.  tests/test_gpu_parity.py compares FFT bins bit for bit.
__device__ __forceinline__ c64 pcx_mul(c64 a, float2 w)
{
    const float2 av = cunpack(a);
    const c64 P = p_mul(a, cpack(w.x, w.x));
    const c64 Q = p_mul(cpack(av.y, av.x), cpack(w.y, w.y));
    return p_fma(Q, cpack(-1.0f, 1.0f), P);
}

// |a|^2 = re*re + im*im (two rounded products, one rounded sum).  The products use
// the packed multiply; the sum is a scalar add.rn.f32 on the two halves (a scalar add
// is never contracted with a packed multiply).
__device__ __forceinline__ float pcx_mag2(c64 a)
{
    const float2 s = cunpack(p_mul(a, a));
    return __fadd_rn(s.x, s.y);
}

// kissfft.hh:137-157 kf_bfly4 (forward), operands already twiddled.
__device__ __forceinline__ void p_bfly4(c64 &a0, c64 &a1, c64 &a2, c64 &a3)
{
    const c64 s5 = p_sub(a0, a2);
    const c64 f0 = p_add(a0, a2);
    const c64 s3 = p_add(a1, a3);
    const c64 s4 = p_sub(a1, a3);
    const float2 u = cunpack(s4);
    const c64 s4r = cpack(u.y, -u.x); // * (-j): folds into an operand modifier
    a2 = p_sub(f0, s3);
    a0 = p_add(f0, s3);
    a1 = p_add(s5, s4r);
    a3 = p_sub(s5, s4r);
}
// kissfft.hh:128-135 kf_bfly2, operand already twiddled.
__device__ __forceinline__ void p_bfly2(c64 &a0, c64 &a1)
{
    const c64 t = a1;
    a1 = p_sub(a0, t);
    a0 = p_add(a0, t);
}

} // namespace lorab200
