// Dumps the fused kernel's work list (vvc-mip-gpu_b200/csrc/mip_work_list.h) for tests/test_work_list.py.
// usage: work_list_dump N0 N1 [w0,w1,.. [v0,v1,..]]  (chunks of the two splits, optional cost shares)
// stdout (binary, little endian): for each half: u32 ntasks, ntasks*32 records (u32 x, u32 y), u32 ncu, ncu * u16 ord2cu;
// then for each split: u32 chunks, for each half: (chunks + 1) * i32 begin, (chunks + 1) * u16 chunk_ord.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "mip_work_list.h"

static std::vector<double> shares(const char* s) {
    std::vector<double> v;
    for (const char* p = s; *p;) {
        char* e;
        v.push_back(strtod(p, &e));
        p = *e == ',' ? e + 1 : e;
        if (e == p && *p) break;
    }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    int nchunks[2] = {atoi(argv[1]), atoi(argv[2])};
    std::vector<double> w[2];
    const double* wp[2] = {nullptr, nullptr};
    for (int sp = 0; sp < 2; ++sp)
        if (argc > 3 + sp && strcmp(argv[3 + sp], "-") != 0) {
            w[sp] = shares(argv[3 + sp]);
            if ((int)w[sp].size() != nchunks[sp]) return 2;
            wp[sp] = w[sp].data();
        }
    static mipb200::WorkList wl;
    if (!mipb200::build_work_list(wl)) { fprintf(stderr, "build_work_list failed\n"); return 1; }
    if (!mipb200::split_work_list(wl, nchunks, wp)) { fprintf(stderr, "split_work_list failed\n"); return 3; }
    auto put32 = [](uint32_t v) { fwrite(&v, 4, 1, stdout); };
    for (int hf = 0; hf < 2; ++hf) {
        put32((uint32_t)wl.wcost[hf].size());
        fwrite(wl.lanes[hf].data(), sizeof(mipb200::LaneRec), wl.lanes[hf].size(), stdout);
        put32((uint32_t)wl.ord_total[hf]);
        fwrite(wl.ord2cu[hf], 2, wl.ord_total[hf], stdout);
    }
    for (int sp = 0; sp < 2; ++sp) {
        put32((uint32_t)wl.chunks[sp]);
        for (int hf = 0; hf < 2; ++hf) {
            fwrite(wl.begin[sp][hf], 4, wl.chunks[sp] + 1, stdout);
            fwrite(wl.chunk_ord[sp][hf], 2, wl.chunks[sp] + 1, stdout);
        }
    }
    return 0;
}
