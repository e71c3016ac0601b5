// Microbenchmark: host->device copy bandwidth of a 64 MiB batch split into chunks over K streams,
// page-locked vs write-combined host memory.  nvcc -O3 -o h2d h2d.cu ; run on the GPU box.
#include <cstdio>
#include <cstring>
#include <chrono>
#include <vector>
#include <cuda_runtime.h>
int main()
{
    const size_t total = 64u << 20;
    void *d; cudaMalloc(&d, total);
    for (int wc = 0; wc < 2; ++wc) {
        void *h;
        cudaHostAlloc(&h, total, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
        memset(h, 1, total);
        for (int K : {1, 2, 3, 4, 8}) {
            for (size_t chunk : {size_t(1) << 20, size_t(4) << 20, size_t(16) << 20, total}) {
                std::vector<cudaStream_t> st(K);
                for (auto &s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
                auto run = [&]() {
                    size_t k = 0;
                    for (size_t off = 0; off < total; off += chunk, ++k)
                        cudaMemcpyAsync((char *)d + off, (char *)h + off, chunk, cudaMemcpyHostToDevice, st[k % K]);
                    for (auto &s : st) cudaStreamSynchronize(s);
                };
                run();
                auto t0 = std::chrono::steady_clock::now();
                for (int r = 0; r < 5; ++r) run();
                auto t1 = std::chrono::steady_clock::now();
                double s = std::chrono::duration<double>(t1 - t0).count() / 5;
                printf("%s streams %d chunk %3zu MiB: %.1f GB/s\n", wc ? "WC    " : "pinned", K, chunk >> 20, total / s / 1e9);
                for (auto &s2 : st) cudaStreamDestroy(s2);
            }
        }
        cudaFreeHost(h);
    }
    return 0;
}
