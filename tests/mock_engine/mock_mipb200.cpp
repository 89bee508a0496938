// mock_mipb200.cpp -- TEST DOUBLE of the C ABI in include/mipb200.h, for CPU-only tests of the CLI's plumbing.
//
// It lives under tests/, is built only by tests/test_cli_mock.py (never by build(), never shipped, never loaded by the
// product) and answers every frame synchronously with the CPU oracle.  Its one purpose: let `pytest -m "not gpu"` run
// csrc/main.cpp end to end -- option handling, worker threads, frame sharding, result rings, the text / decisions / binary
// logs -- on a box without a GPU.  It says nothing about the CUDA engine, whose parity tests are the `-m gpu` ones.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <mutex>
#include <vector>

#include "../../include/mipb200.h"
#include "../../vvc-mip-gpu_b200/csrc/mip_compact.h"

extern "C" {
int mipo_set_bit_depth(int bits);
int mipo_run_frame(const uint16_t* frame, int W, int H, int filter_type, int kernel_idx, int32_t* cost, int32_t* sad, int32_t* satd, int threads);
void mipo_decisions(const int32_t* cost, int nctu, uint8_t* best_mode, int32_t* best_cost);
}

namespace {
thread_local char g_err[256] = "";
// CU type table: modes and offsets, from the same generated header the oracle uses
#include "../../vvc-mip-gpu_b200/csrc/mip_tables.h"

struct Frame {
    int64_t poc;
    std::vector<uint16_t> px;
};
}  // namespace

struct mipb200_engine {
    mipb200_config cfg;
    int n_ctus;
    std::deque<Frame> fifo;
    std::vector<uint16_t> input;
    std::vector<int32_t> cost, sad, satd, best_cost, topk_cost;
    std::vector<uint8_t> best_mode, topk_mode, compact;
};

#define MOCK_API extern "C" __attribute__((visibility("default")))

MOCK_API const char* mipb200_last_error(void) { return g_err; }
MOCK_API const char* mipb200_version(void) { return "mipb200 MOCK (tests only)"; }
MOCK_API int mipb200_num_ctus(int w, int h) { return ((w + 127) / 128) * ((h + 127) / 128); }
MOCK_API int mipb200_device_count(void) { const char* e = getenv("MOCK_GPUS"); return e ? atoi(e) : 1; }
MOCK_API int mipb200_pin_host(void*, size_t) { return 0; }
MOCK_API int mipb200_pin_host_on(int, void*, size_t) { return 0; }
MOCK_API int mipb200_unpin_host(void*) { return 0; }
MOCK_API int mipb200_device_energy_mj(int, unsigned long long*) { strcpy(g_err, "mock: no energy counter"); return MIPB200_ENODEV; }
MOCK_API long long mipb200_kernel_launches(const mipb200_engine*) { return 0; }
MOCK_API int mipb200_in_flight(const mipb200_engine* e) { return (int)e->fifo.size(); }
MOCK_API int mipb200_set_launch_mode(mipb200_engine*, int) { return 0; }

MOCK_API int mipb200_create(mipb200_engine** out, const mipb200_config* cfg) {
    if (cfg->width % 8 || cfg->height % 4 || cfg->device >= mipb200_device_count()) { strcpy(g_err, "mock: bad configuration"); return MIPB200_EINVAL; }
    mipb200_engine* e = new mipb200_engine();
    e->cfg = *cfg;
    if (!e->cfg.bit_depth) e->cfg.bit_depth = 10;
    e->n_ctus = mipb200_num_ctus(cfg->width, cfg->height);
    e->input.resize((size_t)cfg->width * cfg->height);
    *out = e;
    return 0;
}
MOCK_API void mipb200_destroy(mipb200_engine* e) { delete e; }
MOCK_API uint16_t* mipb200_next_input(mipb200_engine* e) { return (int)e->fifo.size() < e->cfg.slots ? e->input.data() : nullptr; }

MOCK_API int mipb200_submit(mipb200_engine* e, const uint16_t* frame, int64_t poc) {
    if ((int)e->fifo.size() >= e->cfg.slots) { strcpy(g_err, "mock: all slots in flight"); return MIPB200_EBUSY; }
    Frame f;
    f.poc = poc;
    f.px.assign(frame, frame + (size_t)e->cfg.width * e->cfg.height);
    e->fifo.push_back(std::move(f));
    return 0;
}

MOCK_API int mipb200_collect(mipb200_engine* e, mipb200_result* r) {
    if (e->fifo.empty()) { strcpy(g_err, "mock: nothing in flight"); return MIPB200_EEMPTY; }
    Frame f = std::move(e->fifo.front());
    e->fifo.pop_front();
    const size_t ncost = (size_t)e->n_ctus * MIP_COSTS_PER_CTU, ncu = (size_t)e->n_ctus * MIP_CUS_PER_CTU;
    e->cost.resize(ncost); e->sad.resize(ncost); e->satd.resize(ncost); e->best_mode.resize(ncu); e->best_cost.resize(ncu);
    static std::mutex mu;   // the oracle's bit depth is a global
    {
        std::lock_guard<std::mutex> lk(mu);
        mipo_set_bit_depth(e->cfg.bit_depth);
        const int rc = mipo_run_frame(f.px.data(), e->cfg.width, e->cfg.height, e->cfg.filter_type, e->cfg.kernel_idx, e->cost.data(), e->sad.data(), e->satd.data(), 2);
        mipo_set_bit_depth(10);
        if (rc) { strcpy(g_err, "mock: oracle failed"); return MIPB200_EINVAL; }
    }
    mipo_decisions(e->cost.data(), e->n_ctus, e->best_mode.data(), e->best_cost.data());
    const int k = e->cfg.top_k > 1 ? e->cfg.top_k : 0;
    if (k) {   // stable selection by (cost, mode)
        e->topk_mode.assign(ncu * k, 0xFF);
        e->topk_cost.assign(ncu * k, -1);
        for (int ctu = 0; ctu < e->n_ctus; ++ctu)
            for (int t = 0; t < MIP_NUM_TYPES; ++t) {
                const mip_cu_type_t& ty = MIP_TYPES[t];
                for (int cu = 0; cu < ty.n; ++cu) {
                    const int32_t* c = e->cost.data() + (size_t)ctu * MIP_COSTS_PER_CTU + ty.cost_off + (size_t)cu * ty.modes;
                    if (c[0] == -1) continue;
                    std::vector<int> order(ty.modes);
                    for (int m = 0; m < ty.modes; ++m) order[m] = m;
                    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return c[a] < c[b]; });
                    const size_t o = ((size_t)ctu * MIP_CUS_PER_CTU + ty.cu_off + cu) * k;
                    for (int j = 0; j < k; ++j) { e->topk_mode[o + j] = (uint8_t)order[j]; e->topk_cost[o + j] = c[order[j]]; }
                }
            }
    }
    const unsigned em = e->cfg.emit;
    r->cost_compact = nullptr;
    if (em & MIPB200_EMIT_COSTS_COMPACT) {
        e->compact.resize((size_t)e->n_ctus * MIP_COMPACT_BYTES_PER_CTU);
        if (mip_compact_pack(e->cost.data(), e->compact.data(), e->n_ctus) != 0) { strcpy(g_err, "mock: a narrow cost does not fit 16 bits"); return MIPB200_EINVAL; }
        r->cost_compact = e->compact.data();
    }
    r->poc = f.poc;
    r->n_ctus = e->n_ctus;
    r->cost = (em & MIPB200_EMIT_COSTS) ? e->cost.data() : nullptr;
    r->sad = (em & MIPB200_EMIT_SAD_SATD) ? e->sad.data() : nullptr;
    r->satd = (em & MIPB200_EMIT_SAD_SATD) ? e->satd.data() : nullptr;
    r->best_mode = (em & MIPB200_EMIT_DECISIONS) ? e->best_mode.data() : nullptr;
    r->best_cost = (em & MIPB200_EMIT_DECISIONS) ? e->best_cost.data() : nullptr;
    r->gpu_ms = 0.f;
    r->top_k = k;
    r->topk_mode = k ? e->topk_mode.data() : nullptr;
    r->topk_cost = k ? e->topk_cost.data() : nullptr;
    return 0;
}
