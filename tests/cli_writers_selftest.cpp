// CPU self-test of the CLI's log formatters (no GPU needed): the buffered, thread-parallel writers of
// vvc-mip-gpu_b200/csrc/main.cpp must produce byte for byte what a plain fprintf loop over the same tables produces
// (the reference's format, main_aux_functions.h:735-798).  Built and run by tests/test_cli.py.
#define main mipb200_cli_main
#include "../vvc-mip-gpu_b200/csrc/main.cpp"
#undef main

#include <random>

static std::string slurp(const char* path) {
    FILE* f = fopen(path, "rb");
    std::string s;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    const int W = 384, H = 256, nCtus = 6, k = 3, nFrames = 3;
    (void)H;
    std::mt19937 rng(12345);
    const size_t ncost = (size_t)nCtus * MIP_COSTS_PER_CTU, ncu = (size_t)nCtus * MIP_CUS_PER_CTU;
    int bad = 0;

    // ---- decisions log
    std::vector<std::vector<uint8_t>> bm(nFrames, std::vector<uint8_t>(ncu * k));
    std::vector<std::vector<int32_t>> bc(nFrames, std::vector<int32_t>(ncu * k));
    for (int p = 0; p < nFrames; ++p)
        for (size_t i = 0; i < ncu * k; ++i) {
            const bool skip = rng() % 17 == 0;
            bm[p][i] = skip ? 255 : (uint8_t)(rng() % 32);
            bc[p][i] = skip ? -1 : (int32_t)(rng() % 40000000);
        }
    {
        const std::string a = dir + "/dec_parallel.csv", b = dir + "/dec_plain.csv";
        FILE* fa = fopen(a.c_str(), "w");
        format_parallel(fa, nFrames, 1024, [&](LogBuf& lb, int poc) { write_decisions(lb, poc, bm[poc].data(), bc[poc].data(), k, 0, nCtus, W); });
        fclose(fa);
        FILE* fb = fopen(b.c_str(), "w");
        for (int poc = 0; poc < nFrames; ++poc)
            for (int ctu = 0; ctu < nCtus; ++ctu)
                for (int t = 0; t < MIP_NUM_TYPES; ++t) {
                    const mip_cu_type_t& ty = MIP_TYPES[t];
                    for (int cu = 0; cu < ty.n; ++cu) {
                        const size_t i = ((size_t)ctu * MIP_CUS_PER_CTU + ty.cu_off + cu) * k;
                        fprintf(fb, "%d,%d,%s,%d,%d,%d,%d,%d", poc, ctu, ty.name, ty.w, ty.h, cu, 128 * (ctu % 3) + ty.xs[cu % ty.cols], 128 * (ctu / 3) + ty.ys[cu / ty.cols]);
                        for (int j = 0; j < k; ++j) fprintf(fb, ",%d,%d", bm[poc][i + j], bc[poc][i + j]);
                        fprintf(fb, "\n");
                    }
                }
        fclose(fb);
        if (slurp(a.c_str()) != slurp(b.c_str())) { fprintf(stderr, "decisions log differs\n"); ++bad; }
    }

    // ---- cost log: frame-0 form (no POC, true SAD/SATD), all-frames compat form (POC, zeros)
    std::vector<int32_t> cost(ncost), sad(ncost), satd(ncost);
    for (size_t i = 0; i < ncost; ++i) {
        const bool skip = rng() % 29 == 0;
        cost[i] = skip ? -1 : (int32_t)(rng() % 30000000);
        sad[i] = skip ? -1 : (int32_t)(rng() % 5000000);
        satd[i] = skip ? -1 : (int32_t)(rng() % 30000000);
    }
    for (int variant = 0; variant < 2; ++variant) {
        const bool withPoc = variant == 1, compat = variant == 1;
        const long poc = variant == 1 ? 7 : 0;
        const std::string a = dir + "/cost_parallel" + std::to_string(variant) + ".csv", b = dir + "/cost_plain" + std::to_string(variant) + ".csv";
        FILE* fa = fopen(a.c_str(), "w");
        format_parallel(fa, nCtus, 4096, [&](LogBuf& lb, int ctu) { write_frame_log(lb, poc, withPoc, cost.data(), compat ? nullptr : sad.data(), compat ? nullptr : satd.data(), ctu, ctu + 1, W, compat); });
        fclose(fa);
        FILE* fb = fopen(b.c_str(), "w");
        for (int ctu = 0; ctu < nCtus; ++ctu)
            for (int t = 0; t < MIP_NUM_TYPES; ++t) {
                const mip_cu_type_t& ty = MIP_TYPES[t];
                for (int cu = 0; cu < ty.n; ++cu)
                    for (int m = 0; m < ty.modes; ++m) {
                        const size_t i = (size_t)ctu * MIP_COSTS_PER_CTU + ty.cost_off + (size_t)cu * ty.modes + m;
                        if (withPoc) fprintf(fb, "%ld,", poc);
                        fprintf(fb, "%d,%s,%d,%d,%d,%d,%d,%d,%d,%d,%d\n", ctu, ty.name, ty.w, ty.h, cu, 128 * (ctu % 3) + ty.xs[cu % ty.cols],
                                128 * (ctu / 3) + ty.ys[cu / ty.cols], m, compat ? 0 : sad[i], compat ? 0 : satd[i], cost[i]);
                    }
            }
        fclose(fb);
        if (slurp(a.c_str()) != slurp(b.c_str())) { fprintf(stderr, "cost log variant %d differs\n", variant); ++bad; }
    }
    printf("cli writers selftest: %s\n", bad ? "FAILED" : "ok");
    return bad ? 1 : 0;
}
