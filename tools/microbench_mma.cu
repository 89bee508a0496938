// microbench_mma.cu -- warp-level tensor-core (mma.sync) issue rate on one B200 next to integer ALU work.
//
// Question it answers ("SATD through the tensor cores"): could the 4x4 Hadamard of the SATD
// (16 x 16 matrix of +-1 applied to 16 exact small integers) run as HMMA.16816.F32 / IMMA.16832.S8 while
// the integer pipes do the rest, i.e. how many mma.sync per clock per SM are there, and do they co-issue
// with LOP3/IMAD?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_mma tools/microbench_mma.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define ITERS 2048
#define ACC 4   // independent accumulator sets per warp

__device__ __forceinline__ void hmma16816(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void imma16832(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MODE 0: HMMA only; 1: IMMA only; 2: HMMA + ALUOPS integer ops per HMMA; 3: integer ops only (same count as 2)
template <int MODE, int ALUOPS>
__global__ void __launch_bounds__(256) k_mma(int* out, unsigned a0, unsigned b0) {
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = a0 + threadIdx.x * (i + 1);
    for (int i = 0; i < 2; ++i) b[i] = b0 ^ (threadIdx.x << i);
    float cf[ACC][4];
    int ci[ACC][4];
    int r[8];
    for (int k = 0; k < ACC; ++k)
        for (int i = 0; i < 4; ++i) { cf[k][i] = 0.f; ci[k][i] = 0; }
    for (int i = 0; i < 8; ++i) r[i] = a0 + i;
    const int bb = b0 + threadIdx.x, dd = b0 ^ 0x55;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < ACC; ++k) {
            if (MODE == 0 || MODE == 2) hmma16816(cf[k], a, b);
            if (MODE == 1) imma16832(ci[k], a, b);
            if (MODE == 2 || MODE == 3) {
#pragma unroll
                for (int q = 0; q < ALUOPS; ++q) {
                    if (q & 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[q & 7]) : "r"(bb), "r"(dd));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[q & 7]) : "r"(bb), "r"(dd));
                }
            }
        }
    }
    float sf = 0.f;
    int si = 0;
    for (int k = 0; k < ACC; ++k)
        for (int i = 0; i < 4; ++i) { sf += cf[k][i]; si ^= ci[k][i]; }
    for (int i = 0; i < 8; ++i) si ^= r[i];
    if (sf == 123.456f || si == 0x7fffffff) out[blockIdx.x * blockDim.x + threadIdx.x] = si;
}

// warp shuffles: 8 independent chains
__global__ void __launch_bounds__(256) k_shfl(int* out, int a0) {
    int r[8];
    for (int i = 0; i < 8; ++i) r[i] = a0 + threadIdx.x * (i + 1);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __shfl_xor_sync(0xffffffffu, r[i], 1);
    }
    int s = 0;
    for (int i = 0; i < 8; ++i) s ^= r[i];
    if (s == 0x7fffffff) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static double time_kernel(K launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8;
    int* d_out;
    cudaMalloc(&d_out, (size_t)blocks * 256 * 4);
    const double warps = (double)blocks * 8, clk = clk_khz * 1e3;
    FILE* js = argc > 1 ? fopen(argv[1], "w") : nullptr;
    if (js) fprintf(js, "{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d", p.name, sms, clk_khz / 1000);
    auto report = [&](const char* name, double ms, double mma_per_warp, double alu_per_warp) {
        const double mma = warps * mma_per_warp, alu = warps * alu_per_warp * 32;
        const double mma_per_clk_sm = mma / (ms * 1e-3) / clk / sms;
        printf("%-28s %8.3f ms  mma/clk/SM %6.3f  (f16 dense %7.1f TFLOP/s)  int lane-ops %6.2f Tops/s\n", name, ms, mma_per_clk_sm,
               mma * 4096 / (ms * 1e-3) * 1e-12, alu / (ms * 1e-3) * 1e-12);
        if (js) fprintf(js, ", \"%s\": {\"ms\": %.4f, \"mma_per_clk_per_sm\": %.4f, \"int_tops\": %.3f}", name, ms, mma_per_clk_sm, alu / (ms * 1e-3) * 1e-12);
    };
    const double n = (double)ITERS * ACC;
    report("hmma16816_f32", time_kernel([&] { k_mma<0, 0><<<blocks, 256>>>(d_out, 1, 2); }), n, 0);
    report("imma16832_s8 (k32)", time_kernel([&] { k_mma<1, 0><<<blocks, 256>>>(d_out, 1, 2); }), n, 0);
    report("int_only_8_per_slot", time_kernel([&] { k_mma<3, 8><<<blocks, 256>>>(d_out, 1, 2); }), 0, n * 8);
    report("hmma+8int", time_kernel([&] { k_mma<2, 8><<<blocks, 256>>>(d_out, 1, 2); }), n, n * 8);
    report("int_only_16_per_slot", time_kernel([&] { k_mma<3, 16><<<blocks, 256>>>(d_out, 1, 2); }), 0, n * 16);
    report("hmma+16int", time_kernel([&] { k_mma<2, 16><<<blocks, 256>>>(d_out, 1, 2); }), n, n * 16);
    report("hmma+32int", time_kernel([&] { k_mma<2, 32><<<blocks, 256>>>(d_out, 1, 2); }), n, n * 32);
    report("int_only_32_per_slot", time_kernel([&] { k_mma<3, 32><<<blocks, 256>>>(d_out, 1, 2); }), 0, n * 32);
    {
        const double ms = time_kernel([&] { k_shfl<<<blocks, 256>>>(d_out, 1); });
        const double shfl = warps * ITERS * 8;
        printf("%-28s %8.3f ms  warp-shfl/clk/SM %6.3f\n", "shfl_xor", ms, shfl / (ms * 1e-3) / clk / sms);
        if (js) fprintf(js, ", \"shfl_xor\": {\"ms\": %.4f, \"warp_shfl_per_clk_per_sm\": %.4f}", ms, shfl / (ms * 1e-3) / clk / sms);
    }
    if (js) { fprintf(js, "}\n"); fclose(js); }
    cudaFree(d_out);
    return 0;
}
