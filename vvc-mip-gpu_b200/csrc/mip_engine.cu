// mip_engine.cu -- C ABI (include/mipb200.h) on top of the kernels: pinned host rings, one
// CUDA stream per slot so that H2D of frame f+1, the kernels of frame f and D2H of frame f-1
// overlap, FIFO collection.  Replaces the reference's OpenCL host plumbing (main.cpp:87-315,
// 408-457, 560-598, 678-1250; main_aux_functions.h:585-630).  No CPU fallback exists.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost a pointer test unless a profiler (nsys / ncu --nvtx) is attached

#include "../../include/mipb200.h"
#include "mip_compact.h"
#include "mip_kernels.h"
#include "mip_tables.h"

#define MIPB200_API extern "C" __attribute__((visibility("default")))

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(call)                                                                         \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail(MIPB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_done = nullptr;
    uint16_t* h_frame = nullptr;  // pinned
    uint16_t* d_frame = nullptr;
    int32_t *d_cost = nullptr, *d_sad = nullptr, *d_satd = nullptr, *d_best_cost = nullptr;
    uint8_t* d_best_mode = nullptr;
    int32_t *h_cost = nullptr, *h_sad = nullptr, *h_satd = nullptr, *h_best_cost = nullptr;  // pinned
    uint8_t* h_best_mode = nullptr;
    uint8_t *d_topk_mode = nullptr, *h_topk_mode = nullptr;   // top_k > 1 only
    int32_t *d_topk_cost = nullptr, *h_topk_cost = nullptr;
    int64_t poc = 0;
    bool busy = false;
};

std::mutex g_init_mutex;
bool g_dev_init[64] = {false};

// Every exported function that touches the device switches to the engine's GPU and puts the caller's current device
// back on return: a host application (torch, another engine) must not find its thread's device changed under it.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) err = cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Host ranges page-locked through mipb200_pin_host(): submit recognises them without asking the driver.
struct PinnedRange { const char* lo; const char* hi; };
std::mutex g_pin_mutex;
std::vector<PinnedRange> g_pinned;

bool in_pinned_registry(const void* p, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_pin_mutex);
    const char* c = static_cast<const char*>(p);
    for (const PinnedRange& r : g_pinned)
        if (c >= r.lo && c + bytes <= r.hi) return true;
    return false;
}

struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

}  // namespace

struct mipb200_engine {
    mipb200_config cfg;
    int n_ctus = 0;
    size_t frame_bytes = 0, cost_bytes = 0, cu_bytes4 = 0, cu_bytes1 = 0;
    // cfg.slots frames may be in flight; the ring has ONE MORE slot than that, so that the slot whose results the last
    // mipb200_collect() exposed is never the one the next mipb200_submit() writes into: result pointers stay valid until
    // the next collect, as include/mipb200.h promises.
    std::vector<Slot> slots;
    int head = 0;   // next slot to submit into
    int tail = 0;   // oldest slot in flight
    int in_flight = 0;
    long long launches = 0;
    cudaStream_t aux_stream = nullptr;
    // Read-backs of all slots go through ONE stream, in submission (= collection) order.  On the slots' own streams the
    // copy engine serves whichever frame's kernel happens to finish first, the oldest frame's table can come back last,
    // FIFO collection then waits for it while every slot sits idle, and the pipeline degenerates into bursts of `slots`
    // frames (measured: 850 frames/s with full tables at 1080p; in order: the link rate with the same 3 slots).
    cudaStream_t d2h_stream = nullptr;
    mipb200::FilterParams fp;         // fused low-pass filter of this configuration
    int launch_mode = MIPB200_LAUNCH_AUTO;
    FILE* trace = nullptr;            // MIPB200_TRACE=<file>: per-frame device timeline (debugging aid, not API)
    cudaEvent_t ev_base = nullptr;
};

MIPB200_API const char* mipb200_last_error(void) { return g_err; }
MIPB200_API const char* mipb200_version(void) { return "mipb200 0.2 (sm_100a)"; }
MIPB200_API int mipb200_num_ctus(int w, int h) { return ((w + 127) / 128) * ((h + 127) / 128); }

static cudaError_t device_count(int* n) {
    cudaError_t ce = cudaGetDeviceCount(n);
    for (int attempt = 0; attempt < 10 && (ce == cudaErrorInitializationError || ce == cudaErrorDevicesUnavailable); ++attempt) {
        // seen on a box where the previous process was still tearing its context down: transient, a short wait cures it
        cudaGetLastError();
        usleep(200 * 1000);
        ce = cudaGetDeviceCount(n);
    }
    return ce;
}

MIPB200_API int mipb200_device_count(void) {
    int n = 0;
    const cudaError_t ce = device_count(&n);
    if (ce == cudaErrorNoDevice || ce == cudaErrorInsufficientDriver) { cudaGetLastError(); return 0; }
    if (ce != cudaSuccess) return fail(MIPB200_ECUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(ce));
    return n;
}

static int check_cfg(const mipb200_config* c) {
    if (!c) return fail(MIPB200_EINVAL, "config is NULL");
    if (c->width < 8 || c->height < 4 || c->width % 8 != 0 || c->height % 4 != 0)
        return fail(MIPB200_EINVAL, "unsupported resolution %dx%d (need width %% 8 == 0, height %% 4 == 0)", c->width, c->height);
    if ((long long)mipb200_num_ctus(c->width, c->height) * MIP_COSTS_PER_CTU > 0xffffffffll)
        return fail(MIPB200_EINVAL, "frame %dx%d has more than 43 897 CTUs (the cost index is 32-bit)", c->width, c->height);
    if (c->filter_type < 0 || c->filter_type > 8) return fail(MIPB200_EINVAL, "filter_type %d out of range 0..8", c->filter_type);
    if (c->filter_type > 0) {
        const int nk = c->filter_type >= 5 ? 3 : 5;
        if (c->kernel_idx < 0 || c->kernel_idx >= nk)
            return fail(MIPB200_EINVAL, "kernel_idx %d out of range 0..%d for filter %s", c->kernel_idx, nk - 1,
                        MIP_FILTER_NAMES[c->filter_type - 1]);
    }
    if (c->slots < 1 || c->slots > 16) return fail(MIPB200_EINVAL, "slots %d out of range 1..16", c->slots);
    if (c->emit == 0 || (c->emit & ~15u)) return fail(MIPB200_EINVAL, "emit mask 0x%x invalid", c->emit);
    if (c->emit & MIPB200_EMIT_COSTS_COMPACT) {
        if (c->emit & (MIPB200_EMIT_COSTS | MIPB200_EMIT_SAD_SATD)) return fail(MIPB200_EINVAL, "MIPB200_EMIT_COSTS_COMPACT replaces MIPB200_EMIT_COSTS and excludes MIPB200_EMIT_SAD_SATD");
        if (c->bit_depth == 12) return fail(MIPB200_EINVAL, "the compact cost table holds 16-bit entries: bit_depth 12 needs MIPB200_EMIT_COSTS");
        if (c->top_k > 1) return fail(MIPB200_EINVAL, "top_k %d reads the int32 table: use MIPB200_EMIT_COSTS", c->top_k);
    }
    if (c->top_k < 0 || c->top_k > MIPB200_TOPK_MAX) return fail(MIPB200_EINVAL, "top_k %d out of range 0..%d", c->top_k, MIPB200_TOPK_MAX);
    if (c->bit_depth != 0 && c->bit_depth != 8 && c->bit_depth != 10 && c->bit_depth != 12)
        return fail(MIPB200_EINVAL, "bit_depth %d not one of 0 (= 10), 8, 10, 12", c->bit_depth);
    if (c->top_k > 1 && !(c->emit & MIPB200_EMIT_DECISIONS)) return fail(MIPB200_EINVAL, "top_k %d needs MIPB200_EMIT_DECISIONS", c->top_k);
    return MIPB200_OK;
}

static void free_slot(Slot& s) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFreeHost(s.h_frame); cudaFreeHost(s.h_cost); cudaFreeHost(s.h_sad); cudaFreeHost(s.h_satd);
    cudaFreeHost(s.h_best_cost); cudaFreeHost(s.h_topk_cost);      // the mode arrays live in the same allocations
    cudaFree(s.d_topk_cost);
    cudaFree(s.d_frame); cudaFree(s.d_cost); cudaFree(s.d_sad); cudaFree(s.d_satd);
    cudaFree(s.d_best_cost);
    if (s.ev_start) cudaEventDestroy(s.ev_start);
    if (s.ev_k0) cudaEventDestroy(s.ev_k0);
    if (s.ev_k1) cudaEventDestroy(s.ev_k1);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}

MIPB200_API void mipb200_destroy(mipb200_engine* e) {
    if (!e) return;
    DeviceGuard dg(e->cfg.device);
    if (e->d2h_stream) cudaStreamSynchronize(e->d2h_stream);
    for (auto& s : e->slots) free_slot(s);
    if (e->d2h_stream) { cudaStreamSynchronize(e->d2h_stream); cudaStreamDestroy(e->d2h_stream); }
    if (e->aux_stream) { cudaStreamSynchronize(e->aux_stream); cudaStreamDestroy(e->aux_stream); }
    if (e->ev_base) cudaEventDestroy(e->ev_base);
    if (e->trace) fclose(e->trace);
    delete e;
}

MIPB200_API int mipb200_create(mipb200_engine** out, const mipb200_config* cfg) {
    if (!out) return fail(MIPB200_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = check_cfg(cfg);
    if (rc) return rc;
    int ndev = 0;
    const cudaError_t ce = device_count(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(MIPB200_ENODEV, "no CUDA device: %s (this engine has no CPU fallback)", cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(MIPB200_ENODEV, "device %d not in 0..%d", cfg->device, ndev - 1);
    if (cfg->device >= (int)(sizeof(g_dev_init) / sizeof(g_dev_init[0]))) return fail(MIPB200_ENODEV, "device %d: at most %d devices are supported", cfg->device, (int)(sizeof(g_dev_init) / sizeof(g_dev_init[0])));
    DeviceGuard dg(cfg->device);
    CU_TRY(dg.err);
    {
        std::lock_guard<std::mutex> lk(g_init_mutex);
        if (!g_dev_init[cfg->device]) {
            // Chunks per CTU half (= CTAs per half), two splits (see kernels_init): with frames overlapping on the slot
            // streams 3 equal shares are the throughput optimum (0.435 ms per 1080p frame; 4 chunks: 0.449); a lone frame
            // wants its last CTAs short: shares 4:3:2:1 (0.48 ms instead of 0.52).  Tuning knobs, not API:
            // MIPB200_CHUNK_WEIGHTS=a,b,c,.. / MIPB200_CHUNK_WEIGHTS_LONE=a,b,c,.. (relative cost shares in launch order).
            static const double kEqual3[3] = {1, 1, 1}, kLone[4] = {4, 3, 2, 1};
            double wbuf[2][64];
            int nchunks[2] = {3, 4};
            const double* weights[2] = {kEqual3, kLone};
            const char* names[2] = {"MIPB200_CHUNK_WEIGHTS", "MIPB200_CHUNK_WEIGHTS_LONE"};
            for (int sp = 0; sp < 2; ++sp) {
                const char* ew = getenv(names[sp]);
                if (!ew) continue;
                int n = 0;
                for (const char* p = ew; *p && n < 64;) {
                    char* end = nullptr;
                    const double w = strtod(p, &end);
                    if (end == p || !(w > 0) || (*end && *end != ','))
                        return fail(MIPB200_EINVAL, "%s=\"%s\": need positive numbers separated by commas", names[sp], ew);
                    wbuf[sp][n++] = w;
                    p = *end == ',' ? end + 1 : end;
                }
                if (n < 2) return fail(MIPB200_EINVAL, "%s=\"%s\": need at least two chunks (one chunk would hold more CUs than the kernel's decision table)", names[sp], ew);
                nchunks[sp] = n;
                weights[sp] = wbuf[sp];
            }
            const cudaError_t ke = mipb200::kernels_init(nchunks, weights);
            if (ke == cudaErrorInvalidValue)
                return fail(MIPB200_EINVAL, "a chunk split puts more than 2048 CUs into one chunk; use more chunks or more even weights");
            CU_TRY(ke);
            g_dev_init[cfg->device] = true;
        }
    }
    mipb200_engine* e = new mipb200_engine();
    e->cfg = *cfg;
    if (e->cfg.bit_depth == 0) e->cfg.bit_depth = 10;
    e->n_ctus = mipb200_num_ctus(cfg->width, cfg->height);
    e->frame_bytes = (size_t)cfg->width * cfg->height * sizeof(uint16_t);
    e->cost_bytes = (size_t)e->n_ctus * MIP_COSTS_PER_CTU * sizeof(int32_t);
    e->cu_bytes4 = (size_t)e->n_ctus * MIP_CUS_PER_CTU * sizeof(int32_t);
    e->cu_bytes1 = (size_t)e->n_ctus * MIP_CUS_PER_CTU;
    e->slots.resize(cfg->slots + 1);
    if (mipb200::make_filter_params(cfg->filter_type, cfg->kernel_idx, e->cfg.bit_depth, &e->fp) != cudaSuccess) {
        delete e;
        return fail(MIPB200_EINVAL, "filter parameters of filter_type %d kernel_idx %d failed their exactness check", cfg->filter_type, cfg->kernel_idx);
    }
    const bool wk = cfg->emit & MIPB200_EMIT_COSTS_COMPACT;
    if (wk) e->cost_bytes = (size_t)e->n_ctus * MIP_COMPACT_BYTES_PER_CTU;      // the one table of this engine is the compact one
    const bool wc = (cfg->emit & MIPB200_EMIT_COSTS) || wk, ws = cfg->emit & MIPB200_EMIT_SAD_SATD, wd = cfg->emit & MIPB200_EMIT_DECISIONS;
    const int tk = cfg->top_k > 1 ? cfg->top_k : 0;
#define E_TRY(call)                                                                                     \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            fail(MIPB200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(_e));                        \
            mipb200_destroy(e);                                                                         \
            return _e == cudaErrorMemoryAllocation ? MIPB200_ENOMEM : MIPB200_ECUDA;                    \
        }                                                                                               \
    } while (0)
    E_TRY(cudaStreamCreateWithFlags(&e->aux_stream, cudaStreamNonBlocking));
    E_TRY(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
    for (auto& s : e->slots) {
        E_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        E_TRY(cudaEventCreate(&s.ev_start));
        E_TRY(cudaEventCreate(&s.ev_k0));
        E_TRY(cudaEventCreate(&s.ev_k1));
        E_TRY(cudaEventCreate(&s.ev_done));
        E_TRY(cudaMalloc((void**)&s.d_frame, e->frame_bytes));   // the pinned staging frame is allocated on first use (slot_staging)
        if (wc || tk) {   // decisions-only engines never materialise the 97840-entry table; a top-k shortlist reads it
            E_TRY(cudaMalloc((void**)&s.d_cost, e->cost_bytes));
            if (wc) E_TRY(cudaHostAlloc((void**)&s.h_cost, e->cost_bytes, cudaHostAllocDefault));
        }
        if (tk) {
            E_TRY(cudaMalloc((void**)&s.d_topk_cost, (e->cu_bytes4 + e->cu_bytes1) * tk));
            E_TRY(cudaHostAlloc((void**)&s.h_topk_cost, (e->cu_bytes4 + e->cu_bytes1) * tk, cudaHostAllocDefault));
            s.d_topk_mode = reinterpret_cast<uint8_t*>(s.d_topk_cost) + e->cu_bytes4 * tk;
            s.h_topk_mode = reinterpret_cast<uint8_t*>(s.h_topk_cost) + e->cu_bytes4 * tk;
        }
        if (ws) {
            E_TRY(cudaMalloc((void**)&s.d_sad, e->cost_bytes));
            E_TRY(cudaMalloc((void**)&s.d_satd, e->cost_bytes));
            E_TRY(cudaHostAlloc((void**)&s.h_sad, e->cost_bytes, cudaHostAllocDefault));
            E_TRY(cudaHostAlloc((void**)&s.h_satd, e->cost_bytes, cudaHostAllocDefault));
        }
        if (wd) {   // best costs and best modes share one buffer (int32 part first): one read-back copy per frame
            E_TRY(cudaMalloc((void**)&s.d_best_cost, e->cu_bytes4 + e->cu_bytes1));
            E_TRY(cudaHostAlloc((void**)&s.h_best_cost, e->cu_bytes4 + e->cu_bytes1, cudaHostAllocDefault));
            s.d_best_mode = reinterpret_cast<uint8_t*>(s.d_best_cost) + e->cu_bytes4;
            s.h_best_mode = reinterpret_cast<uint8_t*>(s.h_best_cost) + e->cu_bytes4;
        }
    }
#undef E_TRY
    if (const char* tr = getenv("MIPB200_TRACE")) {
        // one line per collected frame: poc, then ms since engine creation of: upload start, kernel start, kernel end, results on the host
        char name[512];
        snprintf(name, sizeof(name), "%s.gpu%d", tr, cfg->device);
        e->trace = fopen(name, "a");
        if (e->trace && cudaEventCreate(&e->ev_base) == cudaSuccess) {
            cudaEventRecord(e->ev_base, e->aux_stream);
            cudaEventSynchronize(e->ev_base);
            fprintf(e->trace, "# poc h2d_start_ms kernel_start_ms kernel_end_ms d2h_done_ms (%dx%d, slots %d, emit %u)\n", cfg->width, cfg->height, cfg->slots, cfg->emit);
        }
    }
    *out = e;
    return MIPB200_OK;
}

MIPB200_API int mipb200_in_flight(const mipb200_engine* e) { return e ? e->in_flight : 0; }

MIPB200_API int mipb200_set_launch_mode(mipb200_engine* e, int mode) {
    if (!e) return fail(MIPB200_EINVAL, "engine is NULL");
    if (mode != MIPB200_LAUNCH_AUTO && mode != MIPB200_LAUNCH_THROUGHPUT && mode != MIPB200_LAUNCH_LATENCY) return fail(MIPB200_EINVAL, "launch mode %d unknown", mode);
    e->launch_mode = mode;
    return MIPB200_OK;
}
MIPB200_API long long mipb200_kernel_launches(const mipb200_engine* e) { return e ? e->launches : 0; }

// pinned staging frame of a slot: only callers that hand over pageable memory or fill mipb200_next_input() need one
static uint16_t* slot_staging(mipb200_engine* e, Slot& s) {
    if (!s.h_frame && cudaHostAlloc((void**)&s.h_frame, e->frame_bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        s.h_frame = nullptr;
    }
    return s.h_frame;
}

MIPB200_API uint16_t* mipb200_next_input(mipb200_engine* e) {
    if (!e || e->in_flight >= e->cfg.slots) return nullptr;
    DeviceGuard dg(e->cfg.device);
    return slot_staging(e, e->slots[e->head]);
}

// one fused kernel per frame: (filter +) boundaries + prediction + costs + per-CU argmin; counts launches
static int enqueue_kernels(mipb200_engine* e, const uint16_t* d_frame, int32_t* d_cost, int32_t* d_sad,
                           int32_t* d_satd, uint8_t* d_bm, int32_t* d_bc, bool lone, cudaStream_t st) {
    const mipb200_config& c = e->cfg;
    CU_TRY(mipb200::launch_costs(d_frame, c.width, c.height, c.bit_depth, e->fp, d_cost, d_sad, d_satd, d_bm, d_bc, lone, (c.emit & MIPB200_EMIT_COSTS_COMPACT) != 0, st));
    e->launches++;
    return MIPB200_OK;
}

MIPB200_API int mipb200_submit(mipb200_engine* e, const uint16_t* frame, int64_t poc) {
    if (!e || !frame) return fail(MIPB200_EINVAL, "engine or frame is NULL");
    if (e->in_flight >= e->cfg.slots) return fail(MIPB200_EBUSY, "all %d slots in flight; collect first", e->cfg.slots);
    NvtxRange nv("mipb200_submit");
    DeviceGuard dg(e->cfg.device);
    CU_TRY(dg.err);
    Slot& s = e->slots[e->head];
    // Source of the H2D DMA: the slot's own pinned buffer, the caller's buffer if that is page-locked already
    // (mipb200_pin_host / cudaHostAlloc / cudaHostRegister / torch pin_memory: no staging copy), else a staging copy of
    // pageable memory.  Device and managed pointers are refused: this is the host path (mipb200_run_device is the other).
    const uint16_t* src = frame;
    if (frame != s.h_frame && !in_pinned_registry(frame, e->frame_bytes)) {
        cudaPointerAttributes attr;
        const cudaError_t pe = cudaPointerGetAttributes(&attr, frame);
        if (pe != cudaSuccess) cudaGetLastError();
        if (pe == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged))
            return fail(MIPB200_EINVAL, "mipb200_submit takes a host frame; %s memory goes through mipb200_run_device", attr.type == cudaMemoryTypeDevice ? "device" : "managed");
        if (pe != cudaSuccess || attr.type != cudaMemoryTypeHost) {
            if (!slot_staging(e, s)) return fail(MIPB200_ENOMEM, "cannot allocate the pinned staging frame (%zu bytes)", e->frame_bytes);
            memcpy(s.h_frame, frame, e->frame_bytes);
            src = s.h_frame;
        }
    }
    s.poc = poc;
    if (e->trace) CU_TRY(cudaEventRecord(s.ev_start, s.stream));       // every runtime call in here is paid per frame, on every GPU's thread
    CU_TRY(cudaMemcpyAsync(s.d_frame, src, e->frame_bytes, cudaMemcpyHostToDevice, s.stream));
    CU_TRY(cudaEventRecord(s.ev_k0, s.stream));
    // a frame submitted into an empty pipeline has the GPU to itself (at least at its start): the lone-frame split
    const bool lone = e->launch_mode == MIPB200_LAUNCH_LATENCY || (e->launch_mode == MIPB200_LAUNCH_AUTO && e->in_flight == 0);
    int rc = enqueue_kernels(e, s.d_frame, s.d_cost, s.d_sad, s.d_satd, s.d_best_mode, s.d_best_cost, lone, s.stream);
    if (rc) return rc;
    if (s.d_topk_mode) {
        CU_TRY(mipb200::launch_topk(s.d_cost, e->n_ctus, e->cfg.top_k, s.d_topk_mode, s.d_topk_cost, s.stream));
        e->launches++;
    }
    CU_TRY(cudaEventRecord(s.ev_k1, s.stream));
    cudaStream_t rb = e->d2h_stream;           // in-order read-back: decisions first (small), then the tables
    CU_TRY(cudaStreamWaitEvent(rb, s.ev_k1, 0));
    if (s.h_best_cost) CU_TRY(cudaMemcpyAsync(s.h_best_cost, s.d_best_cost, e->cu_bytes4 + e->cu_bytes1, cudaMemcpyDeviceToHost, rb));
    if (s.h_topk_cost) CU_TRY(cudaMemcpyAsync(s.h_topk_cost, s.d_topk_cost, (e->cu_bytes4 + e->cu_bytes1) * e->cfg.top_k, cudaMemcpyDeviceToHost, rb));
    if (s.h_cost) CU_TRY(cudaMemcpyAsync(s.h_cost, s.d_cost, e->cost_bytes, cudaMemcpyDeviceToHost, rb));
    if (s.h_sad) {
        CU_TRY(cudaMemcpyAsync(s.h_sad, s.d_sad, e->cost_bytes, cudaMemcpyDeviceToHost, rb));
        CU_TRY(cudaMemcpyAsync(s.h_satd, s.d_satd, e->cost_bytes, cudaMemcpyDeviceToHost, rb));
    }
    CU_TRY(cudaEventRecord(s.ev_done, rb));
    s.busy = true;
    e->head = (e->head + 1) % (int)e->slots.size();
    e->in_flight++;
    return MIPB200_OK;
}

MIPB200_API int mipb200_collect(mipb200_engine* e, mipb200_result* out) {
    if (!e || !out) return fail(MIPB200_EINVAL, "engine or result is NULL");
    if (e->in_flight == 0) return fail(MIPB200_EEMPTY, "nothing in flight");
    NvtxRange nv("mipb200_collect");
    DeviceGuard dg(e->cfg.device);
    CU_TRY(dg.err);
    Slot& s = e->slots[e->tail];
    CU_TRY(cudaEventSynchronize(s.ev_done));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    if (e->trace && e->ev_base) {
        float t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        cudaEventElapsedTime(&t0, e->ev_base, s.ev_start); cudaEventElapsedTime(&t1, e->ev_base, s.ev_k0);
        cudaEventElapsedTime(&t2, e->ev_base, s.ev_k1); cudaEventElapsedTime(&t3, e->ev_base, s.ev_done);
        fprintf(e->trace, "%lld %.4f %.4f %.4f %.4f\n", (long long)s.poc, t0, t1, t2, t3);
    }
    out->poc = s.poc;
    out->n_ctus = e->n_ctus;
    const bool wk = e->cfg.emit & MIPB200_EMIT_COSTS_COMPACT;
    out->cost = wk ? nullptr : s.h_cost;
    out->cost_compact = wk ? s.h_cost : nullptr;
    out->sad = s.h_sad;
    out->satd = s.h_satd;
    out->best_mode = s.h_best_mode;
    out->best_cost = s.h_best_cost;
    out->gpu_ms = ms;
    out->top_k = s.h_topk_mode ? e->cfg.top_k : 0;
    out->topk_mode = s.h_topk_mode;
    out->topk_cost = s.h_topk_cost;
    s.busy = false;
    e->tail = (e->tail + 1) % (int)e->slots.size();
    e->in_flight--;
    return MIPB200_OK;
}

MIPB200_API int mipb200_run_device(mipb200_engine* e, const uint16_t* d_frame, int32_t* d_cost, int32_t* d_sad, int32_t* d_satd,
                                   uint8_t* d_best_mode, int32_t* d_best_cost, void* stream) {
    if (!e || !d_frame) return fail(MIPB200_EINVAL, "engine and d_frame are required");
    if (!d_cost && !d_best_mode) return fail(MIPB200_EINVAL, "at least one of d_cost and d_best_mode/d_best_cost is required");
    if ((d_best_mode == nullptr) != (d_best_cost == nullptr)) return fail(MIPB200_EINVAL, "d_best_mode and d_best_cost go together");
    if ((d_sad == nullptr) != (d_satd == nullptr)) return fail(MIPB200_EINVAL, "d_sad and d_satd go together");
    DeviceGuard dg(e->cfg.device);
    CU_TRY(dg.err);
    cudaStream_t st = stream ? (cudaStream_t)stream : e->aux_stream;
    // the caller's streams are its business: unless told otherwise (mipb200_set_launch_mode) launches are assumed to overlap
    return enqueue_kernels(e, d_frame, d_cost, d_sad, d_satd, d_best_mode, d_best_cost, e->launch_mode == MIPB200_LAUNCH_LATENCY, st);
}

MIPB200_API int mipb200_filter_device(mipb200_engine* e, const uint16_t* d_frame, uint16_t* d_out, void* stream) {
    if (!e || !d_frame || !d_out) return fail(MIPB200_EINVAL, "engine, d_frame and d_out are required");
    if (!e->cfg.filter_type) return fail(MIPB200_EINVAL, "engine was created with filter_type 0");
    DeviceGuard dg(e->cfg.device);
    CU_TRY(dg.err);
    cudaStream_t st = stream ? (cudaStream_t)stream : e->aux_stream;
    CU_TRY(mipb200::launch_filter(d_frame, d_out, e->cfg.width, e->cfg.height, e->cfg.filter_type, e->cfg.kernel_idx, st));
    e->launches++;
    return MIPB200_OK;
}

MIPB200_API int mipb200_decide_device(mipb200_engine* e, const int32_t* d_cost, uint8_t* d_best_mode, int32_t* d_best_cost, void* stream) {
    if (!e || !d_cost || !d_best_mode || !d_best_cost) return fail(MIPB200_EINVAL, "NULL argument");
    DeviceGuard dg(e->cfg.device);
    CU_TRY(dg.err);
    cudaStream_t st = stream ? (cudaStream_t)stream : e->aux_stream;
    CU_TRY(mipb200::launch_decide(d_cost, e->n_ctus, d_best_mode, d_best_cost, st));
    e->launches++;
    return MIPB200_OK;
}

MIPB200_API int mipb200_topk_device(mipb200_engine* e, const int32_t* d_cost, int k, uint8_t* d_modes, int32_t* d_costs, void* stream) {
    if (!e || !d_cost || !d_modes || !d_costs) return fail(MIPB200_EINVAL, "NULL argument");
    if (k < 1 || k > MIPB200_TOPK_MAX) return fail(MIPB200_EINVAL, "k %d out of range 1..%d", k, MIPB200_TOPK_MAX);
    DeviceGuard dg(e->cfg.device);
    CU_TRY(dg.err);
    cudaStream_t st = stream ? (cudaStream_t)stream : e->aux_stream;
    CU_TRY(mipb200::launch_topk(d_cost, e->n_ctus, k, d_modes, d_costs, st));
    e->launches++;
    return MIPB200_OK;
}

MIPB200_API size_t mipb200_compact_bytes_per_ctu(void) { return MIP_COMPACT_BYTES_PER_CTU; }

MIPB200_API int mipb200_expand_costs(const void* compact, int n_ctus, int32_t* cost, int threads) {
    if (!compact || !cost || n_ctus < 0) return fail(MIPB200_EINVAL, "compact, cost and a non-negative n_ctus are required");
    const int nth = std::max(1, std::min(threads, std::min(n_ctus, 64)));
    if (nth == 1) { mip_compact_expand(compact, cost, 0, n_ctus); return MIPB200_OK; }
    std::vector<std::thread> th;
    for (int t = 0; t < nth; ++t)
        th.emplace_back([=] { mip_compact_expand(compact, cost, (int)((long long)n_ctus * t / nth), (int)((long long)n_ctus * (t + 1) / nth)); });
    for (auto& x : th) x.join();
    return MIPB200_OK;
}

MIPB200_API int mipb200_pin_host_on(int device, void* ptr, size_t bytes) {
    if (!ptr || !bytes) return fail(MIPB200_EINVAL, "ptr and bytes are required");
    int cur = 0;
    if (device < 0) { CU_TRY(cudaGetDevice(&cur)); device = cur; }
    DeviceGuard dg(device);
    CU_TRY(dg.err);
    CU_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));   // portable: every device's engine may DMA from it
    std::lock_guard<std::mutex> lk(g_pin_mutex);
    g_pinned.push_back({static_cast<const char*>(ptr), static_cast<const char*>(ptr) + bytes});
    return MIPB200_OK;
}

MIPB200_API int mipb200_pin_host(void* ptr, size_t bytes) { return mipb200_pin_host_on(-1, ptr, bytes); }

MIPB200_API int mipb200_unpin_host(void* ptr) {
    if (!ptr) return fail(MIPB200_EINVAL, "ptr is NULL");
    {
        std::lock_guard<std::mutex> lk(g_pin_mutex);
        g_pinned.erase(std::remove_if(g_pinned.begin(), g_pinned.end(), [&](const PinnedRange& r) { return r.lo == static_cast<const char*>(ptr); }), g_pinned.end());
    }
    CU_TRY(cudaHostUnregister(ptr));
    return MIPB200_OK;
}

// NVML through dlopen (libnvidia-ml.so.1 ships with the driver, not with the toolkit): the board's energy counter,
// addressed by PCI bus id so that CUDA_VISIBLE_DEVICES reordering cannot pick the wrong board.
MIPB200_API int mipb200_device_energy_mj(int device, unsigned long long* millijoules) {
    typedef int (*init_fn)(void);
    typedef int (*by_pci_fn)(const char*, void**);
    typedef int (*energy_fn)(void*, unsigned long long*);
    static std::mutex mu;
    static void* lib = nullptr;
    static by_pci_fn by_pci = nullptr;
    static energy_fn energy = nullptr;
    if (!millijoules) return fail(MIPB200_EINVAL, "millijoules is NULL");
    std::lock_guard<std::mutex> lk(mu);
    if (!lib) {
        void* h = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) return fail(MIPB200_ENODEV, "NVML not available: %s", dlerror());
        init_fn init = (init_fn)dlsym(h, "nvmlInit_v2");
        by_pci = (by_pci_fn)dlsym(h, "nvmlDeviceGetHandleByPciBusId_v2");
        energy = (energy_fn)dlsym(h, "nvmlDeviceGetTotalEnergyConsumption");
        if (!init || !by_pci || !energy) { dlclose(h); return fail(MIPB200_ENODEV, "NVML lacks the energy-counter entry points"); }
        const int rc = init();
        if (rc != 0) { dlclose(h); return fail(MIPB200_ENODEV, "nvmlInit_v2 failed with %d", rc); }
        lib = h;
    }
    char bus[32];
    CU_TRY(cudaDeviceGetPCIBusId(bus, sizeof(bus), device));
    void* dev = nullptr;
    int rc = by_pci(bus, &dev);
    if (rc != 0) return fail(MIPB200_ENODEV, "nvmlDeviceGetHandleByPciBusId(%s) failed with %d", bus, rc);
    rc = energy(dev, millijoules);
    if (rc != 0) return fail(MIPB200_ENODEV, "nvmlDeviceGetTotalEnergyConsumption failed with %d", rc);
    return MIPB200_OK;
}

MIPB200_API int mipb200_sync(mipb200_engine* e) {
    if (!e) return fail(MIPB200_EINVAL, "engine is NULL");
    DeviceGuard dg(e->cfg.device);
    CU_TRY(dg.err);
    for (auto& s : e->slots) CU_TRY(cudaStreamSynchronize(s.stream));
    CU_TRY(cudaStreamSynchronize(e->d2h_stream));
    CU_TRY(cudaStreamSynchronize(e->aux_stream));
    return MIPB200_OK;
}
