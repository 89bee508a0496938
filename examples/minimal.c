/* minimal.c -- the C ABI of include/mipb200.h from plain C: three synthetic 1080p frames in flight, per-CU decisions and
 * the cost table out.  This is the loop a host like the reference's main.cpp (main.cpp:678-1250) shrinks to.
 *
 *   gcc -O2 -Iinclude examples/minimal.c -Lvvc-mip-gpu_b200/lib -lmipb200 -Wl,-rpath,$PWD/vvc-mip-gpu_b200/lib -o minimal && ./minimal
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "mipb200.h"

int main(int argc, char** argv) {
    const int W = argc > 2 ? atoi(argv[1]) : 1920, H = argc > 2 ? atoi(argv[2]) : 1080, N = 8;
    mipb200_config cfg = {0};
    cfg.width = W; cfg.height = H; cfg.device = 0;
    cfg.filter_type = 8; cfg.kernel_idx = 2;                     /* filterFrame_2d_float_5x5_quarterCtu, KernelIdx 2 */
    cfg.slots = 3;
    cfg.emit = MIPB200_EMIT_COSTS | MIPB200_EMIT_DECISIONS;
    mipb200_engine* eng = NULL;
    if (mipb200_create(&eng, &cfg) != MIPB200_OK) {
        fprintf(stderr, "mipb200_create: %s\n", mipb200_last_error());
        return 2;
    }
    int submitted = 0, collected = 0;
    long long sum_best = 0, skipped = 0;
    while (collected < N) {
        while (submitted < N && mipb200_in_flight(eng) < cfg.slots) {
            uint16_t* f = mipb200_next_input(eng);               /* fill the slot's pinned buffer in place: no staging copy */
            uint32_t s = 0x9E3779B9u * (uint32_t)(submitted + 1);
            for (long i = 0; i < (long)W * H; ++i) { s = s * 1664525u + 1013904223u; f[i] = (uint16_t)(s >> 22); }   /* 10-bit noise */
            if (mipb200_submit(eng, f, submitted) != MIPB200_OK) { fprintf(stderr, "submit: %s\n", mipb200_last_error()); return 1; }
            ++submitted;
        }
        mipb200_result r;
        if (mipb200_collect(eng, &r) != MIPB200_OK) { fprintf(stderr, "collect: %s\n", mipb200_last_error()); return 1; }
        for (long i = 0; i < (long)r.n_ctus * MIPB200_CUS_PER_CTU; ++i) {
            if (r.best_cost[i] == MIPB200_SKIPPED) ++skipped; else sum_best += r.best_cost[i];
        }
        /* decisions are the argmin of the table */
        const int32_t* c0 = r.cost;                              /* CTU 0, type 0 (64x64), CU 0: 12 modes */
        int bm = 0;
        for (int m = 1; m < 12; ++m) if (c0[m] < c0[bm]) bm = m;
        if (bm != r.best_mode[0] || c0[bm] != r.best_cost[0]) { fprintf(stderr, "frame %lld: decision != argmin of the table\n", (long long)r.poc); return 1; }
        printf("frame %lld: %d CTUs, %.3f ms on the GPU, best mode of CU 0 = %d (cost %d)\n", (long long)r.poc, r.n_ctus, r.gpu_ms, r.best_mode[0], r.best_cost[0]);
        ++collected;
    }
    printf("%d frames, sum of best costs %lld, %lld CUs outside the frame, %lld kernel launches\n", N, sum_best, skipped, mipb200_kernel_launches(eng));
    mipb200_destroy(eng);
    return 0;
}
