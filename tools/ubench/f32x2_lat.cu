// Microbenchmark: dependent-issue latency of packed f32x2 ops (1 warp per SM sub-partition,
// C independent chains per thread).  cycles per op = latency / min(C, latency/issue_interval).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int C, int MODE>
__global__ void k(float *out, float seed)
{
    unsigned long long p[C], ca, cb;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ca) : "f"(seed * 0.5f), "f"(seed * 0.25f));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(cb) : "f"(seed * 0.125f), "f"(seed * 0.75f));
#pragma unroll
    for (int i = 0; i < C; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(seed + i), "f"(seed - i + threadIdx.x));
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < C; ++i) {
                if (MODE == 0) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ca));
                if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cb));
                if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ca), "l"(cb));
                if (MODE == 3) { float x, y; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); x = __fmul_rn(x, seed); asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x), "f"(y)); }
            }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) { float x, y; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y; }
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) out[0] = float(t1 - t0) / (ITERS * 4 * C);
}
template <int C, int MODE> void run(const char *n)
{
    float *o; cudaMalloc(&o, 4096);
    k<C, MODE><<<1, 32>>>(o, 1.0001f); cudaDeviceSynchronize();
    float c; cudaMemcpy(&c, o, 4, cudaMemcpyDeviceToHost);
    printf("%-10s chains %d : %.2f cycles per op -> dependent latency ~%.1f\n", n, C, c, c * C);
    cudaFree(o);
}
int main()
{
    run<1, 0>("FMUL2"); run<2, 0>("FMUL2"); run<4, 0>("FMUL2"); run<8, 0>("FMUL2");
    run<1, 1>("FADD2"); run<2, 1>("FADD2"); run<4, 1>("FADD2"); run<8, 1>("FADD2");
    run<1, 2>("FFMA2"); run<2, 2>("FFMA2"); run<4, 2>("FFMA2"); run<8, 2>("FFMA2");
    run<1, 3>("FMUL"); run<4, 3>("FMUL"); run<8, 3>("FMUL");
    return 0;
}
