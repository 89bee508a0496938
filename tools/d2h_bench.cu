// d2h_bench.cu -- what the host link can carry: concurrent device->host copies of 52.8 MB buffers (the size of a 1080p
// int32 cost table) on 1, 2, 4, .. GPUs of the box, no kernels running.  One host thread and three streams per GPU,
// each stream copying into its own pinned buffer, like the engine's slot ring.  Variants: default pinned memory,
// write-combined pinned memory, and default pinned memory with a concurrent host->device stream of 4.15 MB frames.
// Output: one JSON object per (variant, GPU count) on stdout.
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/d2h_bench tools/d2h_bench.cu -lpthread
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

struct Res { double gbs_d2h = 0, gbs_h2d = 0; };

static void worker(int dev, size_t bytes, unsigned flags, bool with_h2d, double seconds, std::atomic<int>* ready, std::atomic<bool>* go, Res* out) {
    CK(cudaSetDevice(dev));
    constexpr int NS = 3;
    cudaStream_t st[NS], up;
    void *h[NS], *d[NS], *hf, *df;
    const size_t fbytes = 1920 * 1080 * 2;
    for (int i = 0; i < NS; ++i) {
        CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
        CK(cudaHostAlloc(&h[i], bytes, flags));
        CK(cudaMalloc(&d[i], bytes));
        CK(cudaMemset(d[i], i + 1, bytes));
    }
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
    CK(cudaHostAlloc(&hf, fbytes, cudaHostAllocDefault));
    memset(hf, 1, fbytes);
    CK(cudaMalloc(&df, fbytes));
    CK(cudaDeviceSynchronize());
    ready->fetch_add(1);
    while (!go->load()) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    long copies = 0, ups = 0;
    for (;;) {
        for (int i = 0; i < NS; ++i) {
            CK(cudaStreamSynchronize(st[i]));           // slot free again
            CK(cudaMemcpyAsync(h[i], d[i], bytes, cudaMemcpyDeviceToHost, st[i]));
            if (with_h2d) { CK(cudaMemcpyAsync(df, hf, fbytes, cudaMemcpyHostToDevice, up)); ++ups; }
            ++copies;
        }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) break;
    }
    for (int i = 0; i < NS; ++i) CK(cudaStreamSynchronize(st[i]));
    CK(cudaStreamSynchronize(up));
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out->gbs_d2h = copies * (double)bytes / dt / 1e9;
    out->gbs_h2d = ups * (double)fbytes / dt / 1e9;
    for (int i = 0; i < NS; ++i) { cudaFreeHost(h[i]); cudaFree(d[i]); cudaStreamDestroy(st[i]); }
    cudaFreeHost(hf); cudaFree(df); cudaStreamDestroy(up);
}

int main(int argc, char** argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 2.0;
    const size_t bytes = argc > 2 ? (size_t)atoll(argv[2]) : (size_t)135 * 97840 * 4;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    struct { const char* name; unsigned flags; bool h2d; } variants[] = {
        {"pinned", cudaHostAllocDefault, false}, {"pinned_write_combined", cudaHostAllocWriteCombined, false}, {"pinned_plus_h2d_frames", cudaHostAllocDefault, true}};
    for (auto& v : variants)
        for (int n = 1; n <= ndev; n *= 2) {
            std::vector<Res> res(n);
            std::vector<std::thread> th;
            std::atomic<int> ready{0};
            std::atomic<bool> go{false};
            for (int g = 0; g < n; ++g) th.emplace_back(worker, g, bytes, v.flags, v.h2d, seconds, &ready, &go, &res[g]);
            while (ready.load() < n) std::this_thread::yield();
            go.store(true);
            for (auto& t : th) t.join();
            double tot = 0, toth = 0, mn = 1e30;
            for (auto& r : res) { tot += r.gbs_d2h; toth += r.gbs_h2d; if (r.gbs_d2h < mn) mn = r.gbs_d2h; }
            printf("{\"variant\": \"%s\", \"gpus\": %d, \"buffer_bytes\": %zu, \"d2h_gbs_total\": %.1f, \"d2h_gbs_min_per_gpu\": %.1f, \"h2d_gbs_total\": %.2f, "
                   "\"tables_per_s_total\": %.0f}\n", v.name, n, bytes, tot, mn, toth, tot * 1e9 / bytes);
            fflush(stdout);
        }
    return 0;
}
