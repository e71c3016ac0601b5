// kbench.cu -- kernel-tuning harness: times lorab200_demod_dev of a given liblorab200.so
// (dlopen, so several builds of the library can be compared in one GPU session) the way
// bench.py does -- K launches captured in one CUDA graph, inputs rotating over > 3x L2 of
// distinct buffers, CUDA events on the launching stream -- without Python in the loop, and
// prints a hash of the result records so that variants can be checked for identical output.
//
// Build (here, no GPU needed):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/kbench tools/kbench.cu -ldl
// Run (GPU box):                tools/kbench --lib lora-sdr_b200/liblorab200.so --sf 7 --windows 65536,131072
//
// Not a parity test (tests/ compare against the oracle); not the bench (bench.py is).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../include/lorab200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// window w: up-chirp at symbol s(w) + complex Gaussian noise at 0 dB per sample
__global__ void synth(float2 *iq, unsigned long long n_windows, int sf, uint32_t seed)
{
    const unsigned long long n = 1ull << sf;
    const unsigned long long total = n_windows * n;
    for (unsigned long long g = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; g < total;
         g += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long w = g >> sf;
        const uint32_t i = uint32_t(g & (n - 1));
        const uint32_t s = mix(uint32_t(w) * 2654435761u + seed) & uint32_t(n - 1);
        const double ph = -M_PI * (i + 1.0) + M_PI * double(i) * (i + 1.0) / double(n) + 2.0 * M_PI * double(s) * double(i) / double(n);
        double sn, cs;
        sincos(ph, &sn, &cs);
        const uint32_t h1 = mix(uint32_t(g) ^ seed), h2 = mix(uint32_t(g >> 32) + h1 + 0x9e3779b9u);
        const float u1 = (float(h1 >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = (float(h2 >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float r = sqrtf(-logf(u1));          // sigma^2 = 0.5 per component
        float ns, nc;
        sincosf(6.2831853f * u2, &ns, &nc);
        iq[g] = make_float2(float(cs) + r * nc, float(sn) + r * ns);
    }
}

__global__ void hash_out(const uint32_t *p, size_t n_words, unsigned long long *acc)
{
    unsigned long long h = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x)
        h += (unsigned long long)mix(p[i] + uint32_t(i) * 0x9e3779b9u) * (i | 1);
    atomicAdd(acc, h);
}

struct Api {
    void *h;
    decltype(&lorab200_create) create;
    decltype(&lorab200_destroy) destroy;
    decltype(&lorab200_demod_dev) demod_dev;
    decltype(&lorab200_last_kernel) last_kernel;
    decltype(&lorab200_last_error) last_error;
};

static std::vector<long long> parse_list(const char *s)
{
    std::vector<long long> v;
    while (*s) { v.push_back(atoll(s)); const char *c = strchr(s, ','); if (!c) break; s = c + 1; }
    return v;
}

int main(int argc, char **argv)
{
    const char *lib = "lora-sdr_b200/liblorab200.so";
    std::vector<long long> sfs = {7}, wins = {65536};
    int steps = 400, reps = 5;
    unsigned flags = LORAB200_FLAG_INPUT_STABLE;
    float fine_step = 0.f; bool use_fine = false;
    const char *tag = "";
    bool print_reps = false;
    int idle_ms = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto nxt = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--lib") lib = nxt();
        else if (a == "--sf") sfs = parse_list(nxt());
        else if (a == "--windows") wins = parse_list(nxt());
        else if (a == "--steps") steps = atoi(nxt());
        else if (a == "--reps") reps = atoi(nxt());
        else if (a == "--flags") flags = unsigned(atoi(nxt()));
        else if (a == "--fine-step") { fine_step = float(atof(nxt())); use_fine = true; }
        else if (a == "--tag") tag = nxt();
        else if (a == "--print-reps") print_reps = true;
        else if (a == "--idle-ms") idle_ms = atoi(nxt());
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    Api api;
    api.h = dlopen(lib, RTLD_NOW | RTLD_LOCAL);
    if (!api.h) { fprintf(stderr, "dlopen %s: %s\n", lib, dlerror()); return 2; }
    api.create = (decltype(api.create))dlsym(api.h, "lorab200_create");
    api.destroy = (decltype(api.destroy))dlsym(api.h, "lorab200_destroy");
    api.demod_dev = (decltype(api.demod_dev))dlsym(api.h, "lorab200_demod_dev");
    api.last_kernel = (decltype(api.last_kernel))dlsym(api.h, "lorab200_last_kernel");
    api.last_error = (decltype(api.last_error))dlsym(api.h, "lorab200_last_error");
    CK(cudaSetDevice(0));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    unsigned long long *acc;
    CK(cudaMalloc(&acc, 8));
    const double peak = 6567.4;
    for (long long sf : sfs) {
        lorab200_ctx *ctx = nullptr;
        int rc = api.create(&ctx, 0, int(sf), flags);
        if (rc != 0) { fprintf(stderr, "create sf=%lld failed: %d %s\n", sf, rc, api.last_error(nullptr)); return 2; }
        const size_t n = size_t(1) << sf;
        for (long long nw : wins) {
            const size_t step_bytes = size_t(nw) * n * 8;
            int nbuf = int(std::min<size_t>(64, std::max<size_t>(2, (3ull * 126 * 1024 * 1024 + step_bytes - 1) / step_bytes + 1)));
            std::vector<float2 *> bufs(nbuf);
            for (int b = 0; b < nbuf; ++b) {
                CK(cudaMalloc(&bufs[b], step_bytes));
                synth<<<2048, 256, 0, st>>>(bufs[b], (unsigned long long)nw, int(sf), 0x10AAu + 7919u * b);
            }
            lorab200_result *out;
            CK(cudaMalloc(&out, size_t(nw) * sizeof(lorab200_result)));
            CK(cudaMemsetAsync(out, 0, size_t(nw) * sizeof(lorab200_result), st));
            CK(cudaStreamSynchronize(st));
            lorab200_fine fine = {1234, fine_step};
            auto launch = [&](int i, cudaStream_t s) {
                lorab200_fine f = fine;
                int r = api.demod_dev(ctx, bufs[i % nbuf], size_t(nw), n, LORAB200_TABLE_UP, use_fine ? &f : nullptr, out, nullptr, nullptr, s);
                if (r != 0) { fprintf(stderr, "demod_dev failed: %d %s\n", r, api.last_error(ctx)); exit(2); }
            };
            // result hash of buffer 0 (same data for every library build)
            launch(0, st);
            CK(cudaMemsetAsync(acc, 0, 8, st));
            hash_out<<<256, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(out), size_t(nw) * 4, acc);
            unsigned long long h = 0;
            CK(cudaMemcpyAsync(&h, acc, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (int i = 0; i < 5; ++i) launch(i, st);
            CK(cudaStreamSynchronize(st));
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < steps; ++i) launch(i, st);
            CK(cudaStreamEndCapture(st, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            CK(cudaGraphLaunch(ge, st));
            CK(cudaStreamSynchronize(st));
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            std::vector<float> ms(reps);
            for (int r = 0; r < reps; ++r) {
                if (idle_ms) { struct timespec ts = {idle_ms / 1000, (idle_ms % 1000) * 1000000L}; nanosleep(&ts, nullptr); }
                CK(cudaEventRecord(e0, st));
                CK(cudaGraphLaunch(ge, st));
                CK(cudaEventRecord(e1, st));
                CK(cudaStreamSynchronize(st));
                CK(cudaEventElapsedTime(&ms[r], e0, e1));
            }
            if (print_reps) {
                printf("reps_us:");
                for (int r = 0; r < reps; ++r) printf(" %.2f", ms[r] * 1e3 / steps);
                printf("\n");
            }
            std::sort(ms.begin(), ms.end());
            const double us = ms[reps / 2] * 1e3 / steps, us_best = ms[0] * 1e3 / steps;
            const double gbs = double(nw) * (8.0 * n + 16.0) / (us * 1e-6) / 1e9;
            printf("{\"tag\": \"%s\", \"sf\": %lld, \"windows\": %lld, \"us_per_step\": %.3f, \"us_best\": %.3f, \"msym_s\": %.1f, "
                   "\"gbs\": %.1f, \"frac\": %.4f, \"kernel\": \"%s\", \"hash\": \"%016llx\", \"nbuf\": %d, \"steps\": %d}\n",
                   tag, sf, nw, us, us_best, double(nw) / us, gbs, gbs / peak, api.last_kernel(ctx), h, nbuf, steps);
            fflush(stdout);
            CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
            CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
            CK(cudaFree(out));
            for (auto b : bufs) CK(cudaFree(b));
        }
        api.destroy(ctx);
    }
    return 0;
}
