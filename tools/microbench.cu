// microbench.cu -- per-instruction issue throughput of the integer / fp32 pipes on one B200.
//
// Purpose: (1) the INT32 roofline denominator that MEASURED_PEAKS.json lacks;
// (2) which SASS ops the MIP kernel should be built from (IMAD vs IADD3 vs packed 16x2 vs fp32).
// Every test runs CHAINS independent dependency chains per thread so that latency is hidden;
// the result is lane-operations per second over the whole chip.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu && ./microbench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CHAINS 8
#define ITERS 4096

#define OP_KERNEL(NAME, BODY)                                                                  \
    __global__ void __launch_bounds__(256) k_##NAME(int* out, int a0, int b0) {               \
        int r[CHAINS];                                                                         \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) r[c] = a0 + threadIdx.x * (c + 1); \
        int b = b0 + threadIdx.x, d = b0 ^ 0x55;                                              \
        (void)d;                                                                               \
        for (int it = 0; it < ITERS; ++it) {                                                   \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { BODY; }                       \
        }                                                                                      \
        int s = 0;                                                                             \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) s ^= r[c];                          \
        if (s == 0x7fffffff) out[blockIdx.x * blockDim.x + threadIdx.x] = s;                   \
    }

// one SASS op per BODY unless noted
OP_KERNEL(imad, asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(iadd3, asm volatile("add.s32 %0, %0, %1;" : "+r"(r[c]) : "r"(b)))
OP_KERNEL(lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(shf, asm volatile("shf.r.wrap.b32 %0, %0, %1, 5;" : "+r"(r[c]) : "r"(b)))
OP_KERNEL(iabs, asm volatile("abs.s32 %0, %0;" : "+r"(r[c])); asm volatile("sub.s32 %0, %0, %1;" : "+r"(r[c]) : "r"(b)))  // 2 ops
OP_KERNEL(vabsdiff, asm volatile("sad.s32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(imnmx, asm volatile("max.s32 %0, %0, %1;" : "+r"(r[c]) : "r"(b)))
OP_KERNEL(dp2a, asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(dp4a, asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(viadd16x2, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r[c]) : "r"(b)))
OP_KERNEL(vimnmx16x2, asm volatile("max.s16x2 %0, %0, %1;" : "+r"(r[c]) : "r"(b)))
OP_KERNEL(prmt, asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(r[c]) : "r"(b)))
OP_KERNEL(fadd, asm volatile("add.f32 %0, %0, %1;" : "+r"(r[c]) : "r"(b)))
OP_KERNEL(ffma, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(fabsadd, asm volatile("{.reg .f32 t; abs.f32 t, %1; add.f32 %0, %0, t;}" : "+r"(r[c]) : "r"(b)))
OP_KERNEL(hadd2, asm volatile("add.f16x2 %0, %0, %1;" : "+r"(r[c]) : "r"(b)))
// mixes (2 or 3 ops per BODY)
OP_KERNEL(mix_imad_iadd3, asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(b), "r"(d)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(mix_ffma_lop3, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(b), "r"(d)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(mix_fadd_imad, asm volatile("add.f32 %0, %0, %1;" : "+r"(r[c]) : "r"(b)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(mix_fadd_imad_lop3, asm volatile("add.f32 %0, %0, %1;" : "+r"(r[c]) : "r"(b)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(b), "r"(d)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(mix_dp2a_lop3, asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(b), "r"(d)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(b), "r"(d)))
OP_KERNEL(mix_viadd16_imad, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r[c]) : "r"(b)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(b), "r"(d)))

// which pipe does an op share? pair it with IMAD (fma pipe) and with LOP3 (alu pipe)
#define IMAD_ asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(b), "r"(d))
#define LOP3_ asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(b), "r"(d))
OP_KERNEL(mix_sad_imad, asm volatile("sad.s32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(b), "r"(d)); IMAD_)
OP_KERNEL(mix_sad_lop3, asm volatile("sad.s32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(b), "r"(d)); LOP3_)
OP_KERNEL(mix_abs_imad, asm volatile("abs.s32 %0, %0;" : "+r"(r[c])); IMAD_)
OP_KERNEL(mix_abs_lop3, asm volatile("abs.s32 %0, %0;" : "+r"(r[c])); LOP3_)
OP_KERNEL(mix_shf_imad, asm volatile("shf.r.wrap.b32 %0, %0, %1, 5;" : "+r"(r[c]) : "r"(b)); IMAD_)
OP_KERNEL(mix_dp2a_imad, asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(b), "r"(d)); IMAD_)
OP_KERNEL(mix_viadd16_lop3, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r[c]) : "r"(b)); LOP3_)
OP_KERNEL(mix_vmax16_imad, asm volatile("max.s16x2 %0, %0, %1;" : "+r"(r[c]) : "r"(b)); asm volatile("min.s16x2 %0, %0, %1;" : "+r"(r[c]) : "r"(d)); IMAD_)
OP_KERNEL(mix_max_imad, asm volatile("max.s32 %0, %0, %1;" : "+r"(r[c]) : "r"(b)); asm volatile("min.s32 %0, %0, %1;" : "+r"(r[c]) : "r"(d)); IMAD_)
OP_KERNEL(mix_prmt_imad, asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(r[c]) : "r"(b)); IMAD_)
OP_KERNEL(mix_lea_imad, r[c] = b + (r[c] >> 3); IMAD_)
OP_KERNEL(mix_lea_lop3, r[c] = b + (r[c] >> 3); LOP3_)

typedef void (*kern_t)(int*, int, int);

static double run(kern_t k, int ops_per_body, int sms, int* d_out, const char* name, FILE* js, bool last) {
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<blocks, threads>>>(d_out, 1, 2);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<<<blocks, threads>>>(d_out, 1, 2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)blocks * threads * ITERS * CHAINS * ops_per_body;
    const double tops = ops / (best * 1e-3) / 1e12;
    printf("%-22s %8.3f ms  %7.2f Tops/s  (%.1f lane-ops/clk/SM @1.965GHz)\n", name, best, tops, tops * 1e12 / sms / 1.965e9);
    fprintf(js, "  \"%s\": %.3f%s\n", name, tops, last ? "" : ",");
    return tops;
}

int main(int argc, char** argv) {
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, dev);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int* d_out;
    cudaMalloc(&d_out, 4 << 20);
    const char* path = argc > 1 ? argv[1] : "microbench.json";
    FILE* js = fopen(path, "w");
    fprintf(js, "{\n  \"device\": \"%s\", \"sms\": %d, \"unit\": \"Tera lane-ops/s\",\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
#define RUN(NAME, OPS, LAST) run(k_##NAME, OPS, sms, d_out, #NAME, js, LAST)
    RUN(imad, 1, false); RUN(iadd3, 1, false); RUN(lop3, 1, false); RUN(shf, 1, false); RUN(iabs, 2, false);
    RUN(vabsdiff, 1, false); RUN(imnmx, 1, false); RUN(dp2a, 1, false); RUN(dp4a, 1, false);
    RUN(viadd16x2, 1, false); RUN(vimnmx16x2, 1, false); RUN(prmt, 1, false);
    RUN(fadd, 1, false); RUN(ffma, 1, false); RUN(fabsadd, 1, false); RUN(hadd2, 1, false);
    RUN(mix_imad_iadd3, 2, false); RUN(mix_ffma_lop3, 2, false); RUN(mix_fadd_imad, 2, false);
    RUN(mix_fadd_imad_lop3, 3, false); RUN(mix_dp2a_lop3, 2, false); RUN(mix_viadd16_imad, 2, false);
    RUN(mix_sad_imad, 2, false); RUN(mix_sad_lop3, 2, false); RUN(mix_abs_imad, 2, false); RUN(mix_abs_lop3, 2, false);
    RUN(mix_shf_imad, 2, false); RUN(mix_dp2a_imad, 2, false); RUN(mix_viadd16_lop3, 2, false); RUN(mix_vmax16_imad, 3, false);
    RUN(mix_max_imad, 3, false); RUN(mix_prmt_imad, 2, false); RUN(mix_lea_imad, 2, false); RUN(mix_lea_lop3, 2, true);
    fprintf(js, "}\n");
    fclose(js);
    return 0;
}
