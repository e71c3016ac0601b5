// Microbenchmark: throughput of packed f32x2 ops vs scalar FP32 on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void k(float *out, float seed)
{
    // 8 independent chains per thread
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
    const float c0 = seed * 0.5f, c1 = seed * 0.25f;
    unsigned long long ca, cb;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ca) : "f"(c0), "f"(c1));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(cb) : "f"(c1), "f"(c0));
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { // scalar FMUL + FADD alternating (2 ops per float)
                a[2 * i] = __fmul_rn(a[2 * i], c0); a[2 * i + 1] = __fmul_rn(a[2 * i + 1], c1);
                a[2 * i] = __fadd_rn(a[2 * i], c1); a[2 * i + 1] = __fadd_rn(a[2 * i + 1], c0);
            } else if (MODE == 1) { // packed mul + add
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ca));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cb));
            } else if (MODE == 2) { // packed fma
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ca), "l"(cb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(cb), "l"(ca));
            } else if (MODE == 3) { // scalar FFMA
                a[2 * i] = __fmaf_rn(a[2 * i], c0, c1); a[2 * i + 1] = __fmaf_rn(a[2 * i + 1], c1, c0);
                a[2 * i] = __fmaf_rn(a[2 * i], c1, c0); a[2 * i + 1] = __fmaf_rn(a[2 * i + 1], c0, c1);
            } else if (MODE == 4) { // packed add only
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ca));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cb));
            } else if (MODE == 5) { // packed mul + add with a SCALAR alu op interleaved (lop3) to see co-issue
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ca));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cb));
                a[i] = __int_as_float(__float_as_int(a[i]) ^ (it + i));
                a[8 + i] = __int_as_float(__float_as_int(a[8 + i]) + (it ^ i));
            } else if (MODE == 6) { // scalar mul+add with the same 2 alu ops
                a[2 * i] = __fmul_rn(a[2 * i], c0); a[2 * i + 1] = __fmul_rn(a[2 * i + 1], c1);
                a[2 * i] = __fadd_rn(a[2 * i], c1); a[2 * i + 1] = __fadd_rn(a[2 * i + 1], c0);
                p[i] ^= (unsigned long long)(it + i);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = float(t1 - t0);
}

template <int MODE>
void run(const char *name, int threads)
{
    float *out;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    k<MODE><<<148, threads>>>(out, 1.0001f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(out, 1.0001f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    float cyc; cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
    // float-ops per thread: 16 floats * 2 ops * ITERS
    double flop_lane_ops = 16.0 * 2 * ITERS * threads;      // per SM
    printf("%-34s threads/SM %4d  cycles %.0f  lane-flops/clk/SM %.1f  (%.3f ms, err=%s)\n", name, threads, cyc,
           flop_lane_ops / cyc, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("scalar FMUL+FADD", threads);
        run<1>("packed mul.f32x2+add.f32x2", threads);
        run<4>("packed add.f32x2 only", threads);
        run<3>("scalar FFMA", threads);
        run<2>("packed fma.f32x2", threads);
        run<6>("scalar mul+add + 1 alu / 4 fp", threads);
        run<5>("packed mul+add + 2 alu / 2 packed", threads);
    }
    return 0;
}
