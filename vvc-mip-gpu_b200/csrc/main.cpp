// main.cpp -- command-line host of the B200 MIP engine; drop-in for the reference's ./main.
//
// Same surface as the reference host (main.cpp:43-85 CLI, :364-384 CSV input,
// main_aux_functions.h:735-798 cost log, :908-914 timing block):
//
//   ./mipb200_main -f N -s WxH -o frames.csv [-l prefix] [--DeviceIndex i]
//                  [--FilterType name] [--KernelIdx k]
//
// Long options accept any unique prefix (boost::program_options' default "guessing"), so the
// README's --Filter=... works; values may follow as "--Opt=value" or "--Opt value".
// USE_ALTERNATIVE_SAMPLES is the reference's compile-time switch (main.cpp:10); here it is the
// default of the run-time option --UseAlternativeSamples (0|1).
// Extensions (all off by default): --NumGpus G (frames sharded poc % G over G GPUs, one host
// thread each, no collective), --AllFrames (log every frame with a leading POC column),
// --Compat (print 0 in the SAD/SATD columns like the reference's MAX_PERFORMANCE_DIST build),
// --NoLog (skip the text log), --InputFormat csv|u16|yuv420p|yuv420p10le (binary luma input instead of
// the 2 M stoi() calls per 1080p frame), --DecisionsLog FILE (per-CU best mode + cost of EVERY frame,
// keyed by POC,X,Y,W,H: the table an encoder-side consumer ingests), --TopK k (shortlists in that log),
// --BinaryLog FILE (raw int32 cost tables of every frame), --BitDepth 8|10|12, --Energy (NVML joules per frame),
// --StageStamps 0|1 (the reference's TRACE_POWER stamps).
//
// The device work goes through the C ABI of include/mipb200.h only.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mipb200.h"
#include "mip_tables.h"

#ifndef USE_ALTERNATIVE_SAMPLES
#define USE_ALTERNATIVE_SAMPLES 0
#endif
#ifndef TRACE_POWER
#define TRACE_POWER 1   // the reference ships with TRACE_POWER 1 (main_aux_functions.h:3)
#endif

namespace {

struct Options {
    int deviceIndex = 0;   bool deviceSet = false;
    int nFrames = -1;      bool framesSet = false;
    std::string resolution;
    std::string input;     bool inputSet = false;
    std::string prefix;    bool prefixSet = false;
    std::string filter;    bool filterSet = false;
    int kernelIdx = 0;     bool kernelSet = false;
    int useAlt = USE_ALTERNATIVE_SAMPLES;
    int numGpus = 1;
    int topK = 0;
    int bitDepth = 10;
    int stageStamps = -1;   // -1: follow TRACE_POWER when one GPU is used
    std::string inputFormat = "csv", decisionsLog, binaryLog;
    bool allFrames = false, compat = false, noLog = false, help = false, energy = false;
};

const char* kLongOpts[] = {"help", "DeviceIndex", "FramesToBeEncoded", "Resolution", "OriginalFrames", "OutputPreffix",
                           "FilterType", "KernelIdx", "UseAlternativeSamples", "NumGpus", "AllFrames", "Compat", "NoLog",
                           "InputFormat", "DecisionsLog", "TopK", "Energy", "StageStamps", "BitDepth", "BinaryLog"};
const bool kTakesValue[] = {false, true, true, true, true, true, true, true, true, true, false, false, false, true, true,
                            true, false, true, true, true};
constexpr int kNumOpts = sizeof(kLongOpts) / sizeof(kLongOpts[0]);

void print_help() {
    printf("Allowed options:\n"
           "  -h [ --help ]                  produce help message\n"
           "  --DeviceIndex arg (=0)         Index of the GPU device (CUDA ordinal)\n"
           "  -f [ --FramesToBeEncoded ] arg Number of frames to be processed\n"
           "  -s [ --Resolution ] arg        Resolution of the video, in the format 1920x1080\n"
           "  -o [ --OriginalFrames ] arg    Input file for original frames samples\n"
           "  -l [ --OutputPreffix ] arg     Output files preffix with produced costs\n"
           "  --FilterType arg               Type of smoothing filter\n"
           "  --KernelIdx arg (=0)           Index of the filtering kernel used to define the coefficients\n"
           "  --UseAlternativeSamples arg    0|1, run-time form of the USE_ALTERNATIVE_SAMPLES macro\n"
           "  --NumGpus arg (=1)             shard frames over this many GPUs\n"
           "  --AllFrames --Compat --NoLog   log every frame / zero SAD,SATD columns / no text log\n"
           "  --InputFormat arg (=csv)       csv | u16 (raw little-endian luma) | yuv420p | yuv420p10le\n"
           "  --DecisionsLog arg             write POC,CTU,cuSizeName,W,H,CU,X,Y,BestMode,BestCost for every frame\n"
           "  --BinaryLog arg                write every frame's cost table as raw int32 (64-byte header, see INTEGRATION.md)\n"
           "  --TopK arg (=1)                with --DecisionsLog: the k cheapest modes per CU (adds Mode2,Cost2,... columns)\n"
           "  --Energy                       report joules per frame from the board's NVML energy counter\n"
           "  --BitDepth arg (=10)           8 | 10 | 12; 10 is the reference's pipeline (also for 8-bit content taken as is)\n"
           "  --StageStamps arg              0|1, the reference's per-frame START/FINISH stage stamps (TRACE_POWER)\n");
}

// resolves a (possibly abbreviated) long option; -1 unknown, -2 ambiguous
int match_long(const std::string& name) {
    int hit = -1;
    for (int i = 0; i < kNumOpts; ++i) {
        if (name == kLongOpts[i]) return i;
        if (strncmp(kLongOpts[i], name.c_str(), name.size()) == 0) {
            if (hit >= 0) return -2;
            hit = i;
        }
    }
    return hit;
}

bool to_int(const std::string& s, int* v) {
    char* end = nullptr;
    errno = 0;
    long x = strtol(s.c_str(), &end, 10);
    if (errno || end == s.c_str() || *end) return false;
    *v = (int)x;
    return true;
}

bool parse_args(int argc, char** argv, Options& o) {
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        int opt = -1;
        std::string val;
        bool haveVal = false;
        if (a.rfind("--", 0) == 0) {
            std::string name = a.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); haveVal = true; }
            opt = match_long(name);
            if (opt == -1) { fprintf(stderr, "unrecognised option '--%s'\n", name.c_str()); return false; }
            if (opt == -2) { fprintf(stderr, "option '--%s' is ambiguous\n", name.c_str()); return false; }
        } else if (a.size() >= 2 && a[0] == '-') {
            switch (a[1]) {
                case 'h': opt = 0; break;
                case 'f': opt = 2; break;
                case 's': opt = 3; break;
                case 'o': opt = 4; break;
                case 'l': opt = 5; break;
                default: fprintf(stderr, "unrecognised option '%s'\n", a.c_str()); return false;
            }
            if (a.size() > 2) { val = a.substr(a[2] == '=' ? 3 : 2); haveVal = true; }
        } else {
            fprintf(stderr, "too many positional options have been specified on the command line\n");
            return false;
        }
        if (kTakesValue[opt] && !haveVal) {
            if (i + 1 >= argc) { fprintf(stderr, "the required argument for option '--%s' is missing\n", kLongOpts[opt]); return false; }
            val = argv[++i];
        }
        bool ok = true;
        switch (opt) {
            case 0: o.help = true; break;
            case 1: ok = to_int(val, &o.deviceIndex); o.deviceSet = true; break;
            case 2: ok = to_int(val, &o.nFrames); o.framesSet = true; break;
            case 3: o.resolution = val; break;
            case 4: o.input = val; o.inputSet = true; break;
            case 5: o.prefix = val; o.prefixSet = true; break;
            case 6: o.filter = val; o.filterSet = true; break;
            case 7: ok = to_int(val, &o.kernelIdx); o.kernelSet = true; break;
            case 8: ok = to_int(val, &o.useAlt); break;
            case 9: ok = to_int(val, &o.numGpus); break;
            case 10: o.allFrames = true; break;
            case 11: o.compat = true; break;
            case 12: o.noLog = true; break;
            case 13: o.inputFormat = val; break;
            case 14: o.decisionsLog = val; break;
            case 15: ok = to_int(val, &o.topK); break;
            case 16: o.energy = true; break;
            case 17: ok = to_int(val, &o.stageStamps); break;
            case 18: ok = to_int(val, &o.bitDepth); break;
            case 19: o.binaryLog = val; break;
        }
        if (!ok) { fprintf(stderr, "the argument ('%s') for option '--%s' is invalid\n", val.c_str(), kLongOpts[opt]); return false; }
    }
    return true;
}

// parameter echo of checkReportParameters (main_aux_functions.h:113-162)
int report_parameters(const Options& o) {
    int errors = 0;
    printf("-=-= INPUT PARAMETERS =-=-\n");
    if (!o.deviceSet) printf("  Device index not set. Using standard value of %d.\n", o.deviceIndex);
    else printf("  Device Index=%d\n", o.deviceIndex);
    if (!o.prefixSet) printf("  OutputPreffix log file not set. The output will not be written to any file.\n");
    else printf("  OutputPreffix=%s\n", o.prefix.c_str());
    if (o.framesSet) printf("  FramesToBeEncoded=%d\n", o.nFrames);
    else { printf("  [!] ERROR: FramesToBeEncoded not set.\n"); errors++; }
    if (o.inputSet) printf("  InputOriginalFrame=%s\n", o.input.c_str());
    else { printf("  [!] ERROR: Input original frames not set.\n"); errors++; }
    if (o.useAlt) {
        if (o.filterSet) printf("  FilterType=%s\n", o.filter.c_str());
        else { printf("  [!] ERROR: Filter not set.\n"); errors++; }
        if (!o.kernelSet) printf("  KernelIdx not set. Using default value zero.\n");
        else printf("  KernelIdx=%d\n", o.kernelIdx);
    }
    return errors;
}

void print_timestamp(const char* what) {  // main_aux_functions.h:180-189
    if (!TRACE_POWER) return;
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    struct tm* t = localtime(&tv.tv_sec);
    printf("%s @ %02d:%02d:%02d.%03d\n", what, t->tm_hour, t->tm_min, t->tm_sec, (int)(tv.tv_usec / 1000));
}

double now_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// CSV reader (main.cpp:364-384): N*H lines of W comma-separated integers; extra fields ignored.
bool read_frames_csv(const std::string& path, int W, int H, int N, std::vector<uint16_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { perror("error while opening samples files"); return false; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    if (fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); fprintf(stderr, "short read on %s\n", path.c_str()); return false; }
    fclose(f);
    buf[sz] = '\n';
    out.resize((size_t)W * H * N);
    const char* p = buf.data();
    const char* end = p + sz;
    for (long line = 0; line < (long)N * H; ++line) {
        if (p >= end) { fprintf(stderr, "[!] ERROR: %s holds %ld lines, need %ld (%d frames of %d rows)\n", path.c_str(), line, (long)N * H, N, H); return false; }
        uint16_t* dst = out.data() + (size_t)line * W;
        for (int x = 0; x < W; ++x) {
            while (p < end && (*p == ' ' || *p == '\t')) ++p;
            if (p >= end || *p < '0' || *p > '9') {
                fprintf(stderr, "[!] ERROR: line %ld of %s: field %d is not a number (need %d samples per line)\n", line + 1, path.c_str(), x + 1, W);
                return false;
            }
            int v = 0;
            while (*p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
            dst[x] = (uint16_t)v;
            if (*p == ',') ++p;
        }
        while (p < end && *p != '\n') ++p;
        ++p;
    }
    return true;
}

// Binary luma input: u16 = W*H little-endian uint16 per frame; yuv420p = 8-bit planar 4:2:0 (chroma skipped);
// yuv420p10le = 16-bit planar 4:2:0.  Sample values are taken as they are (the pipeline is 10-bit, intra.cl:61).
bool read_frames_binary(const std::string& path, const std::string& fmt, int W, int H, int N, std::vector<uint16_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { perror("error while opening samples files"); return false; }
    const size_t px = (size_t)W * H;
    out.resize(px * N);
    const bool eight = fmt == "yuv420p";
    const size_t chroma = fmt == "u16" ? 0 : px / 2 * (eight ? 1 : 2);
    std::vector<uint8_t> tmp(eight ? px : 0);
    for (int n = 0; n < N; ++n) {
        uint16_t* dst = out.data() + px * n;
        size_t got;
        if (eight) {
            got = fread(tmp.data(), 1, px, f);
            for (size_t i = 0; i < px; ++i) dst[i] = tmp[i];
        } else {
            got = fread(dst, 2, px, f);
        }
        if (got != px) { fprintf(stderr, "[!] ERROR: %s holds fewer than %d frames of %dx%d (%s)\n", path.c_str(), N, W, H, fmt.c_str()); fclose(f); return false; }
        if (chroma && fseek(f, (long)chroma, SEEK_CUR) != 0) { fclose(f); return false; }
    }
    fclose(f);
    return true;
}

// ---- cost log (main_aux_functions.h:735-798)
struct LogBuf {
    FILE* f;                 // nullptr: memory only (the buffer grows; a formatter thread's private buffer)
    std::vector<char> b;
    size_t n = 0;
    explicit LogBuf(FILE* fp, size_t cap = 8u << 20) : f(fp), b(cap) {}
    void flush() { if (n && f) fwrite(b.data(), 1, n, f); if (f) n = 0; }
    void ensure(size_t k) {
        if (n + k <= b.size()) return;
        if (f) flush(); else b.resize(b.size() * 2 + k);
    }
    void put(const char* s, size_t k) { memcpy(b.data() + n, s, k); n += k; }
    void put_int(long v) {
        char t[24];
        int k = 0;
        bool neg = v < 0;
        unsigned long u = neg ? (unsigned long)(-v) : (unsigned long)v;
        do { t[k++] = (char)('0' + u % 10); u /= 10; } while (u);
        if (neg) b[n++] = '-';
        while (k) b[n++] = t[--k];
    }
};

void write_frame_log(LogBuf& lb, long poc, bool withPoc, const int32_t* cost, const int32_t* sad, const int32_t* satd,
                     int ctuBegin, int ctuEnd, int W, bool compat) {
    const int ctuCols = (W + 127) / 128;
    for (int ctu = ctuBegin; ctu < ctuEnd; ++ctu) {
        const int ctuX = 128 * (ctu % ctuCols), ctuY = 128 * (ctu / ctuCols);
        for (int t = 0; t < MIP_NUM_TYPES; ++t) {   // SizeId 2 types, then SizeId 1, then 4x4: the table order
            const mip_cu_type_t& ty = MIP_TYPES[t];
            for (int cu = 0; cu < ty.n; ++cu) {
                char pre[160];
                int pl = 0;
                if (withPoc) pl += snprintf(pre + pl, sizeof(pre) - pl, "%ld,", poc);
                pl += snprintf(pre + pl, sizeof(pre) - pl, "%d,%s,%d,%d,%d,%d,%d,", ctu, ty.name, ty.w, ty.h, cu,
                               ctuX + ty.xs[cu % ty.cols], ctuY + ty.ys[cu / ty.cols]);
                const size_t base = (size_t)ctu * MIP_COSTS_PER_CTU + ty.cost_off + (size_t)cu * ty.modes;
                for (int m = 0; m < ty.modes; ++m) {
                    lb.ensure(256);
                    lb.put(pre, pl);
                    lb.put_int(m); lb.b[lb.n++] = ',';
                    lb.put_int(compat || !sad ? 0 : sad[base + m]); lb.b[lb.n++] = ',';
                    lb.put_int(compat || !satd ? 0 : satd[base + m]); lb.b[lb.n++] = ',';
                    lb.put_int(cost[base + m]); lb.b[lb.n++] = '\n';
                }
            }
        }
    }
}

// per-CU decisions of one frame: POC,CTU,cuSizeName,W,H,CU,X,Y,BestMode,BestCost[,Mode2,Cost2,...] (skipped CUs: 255,-1)
// bm/bc hold k entries per CU in ascending (cost, mode) order.  726 300 lines per 1080p frame: no printf in the CU loop.
void write_decisions(LogBuf& lb, long poc, const uint8_t* bm, const int32_t* bc, int k, int ctuBegin, int ctuEnd, int W) {
    const int ctuCols = (W + 127) / 128;
    for (int ctu = ctuBegin; ctu < ctuEnd; ++ctu) {
        const int ctuX = 128 * (ctu % ctuCols), ctuY = 128 * (ctu / ctuCols);
        for (int t = 0; t < MIP_NUM_TYPES; ++t) {
            const mip_cu_type_t& ty = MIP_TYPES[t];
            char mid[96];                                   // ",<type name>,<w>,<h>," is the same for every CU of the type
            const int ml = snprintf(mid, sizeof(mid), ",%s,%d,%d,", ty.name, ty.w, ty.h);
            for (int cu = 0; cu < ty.n; ++cu) {
                const size_t i = ((size_t)ctu * MIP_CUS_PER_CTU + ty.cu_off + cu) * k;
                lb.ensure(512);
                lb.put_int(poc); lb.b[lb.n++] = ',';
                lb.put_int(ctu);
                lb.put(mid, ml);
                lb.put_int(cu); lb.b[lb.n++] = ',';
                lb.put_int(ctuX + ty.xs[cu % ty.cols]); lb.b[lb.n++] = ',';
                lb.put_int(ctuY + ty.ys[cu / ty.cols]);
                for (int j = 0; j < k; ++j) {
                    lb.b[lb.n++] = ',';
                    lb.put_int(bm[i + j]); lb.b[lb.n++] = ',';
                    lb.put_int(bc[i + j]);
                }
                lb.b[lb.n++] = '\n';
            }
        }
    }
}

// Runs fmt(buffer, item) for item = 0 .. n-1 on the host threads, each into a private memory buffer, and writes the
// buffers to f in item order (items = CTUs of the cost log, frames of the decisions log).
template <class Fmt>
void format_parallel(FILE* f, int n, size_t bufBytes, Fmt fmt) {
    const int nth = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<LogBuf> parts;
    for (int t = 0; t < std::min(nth, n); ++t) parts.emplace_back(nullptr, bufBytes);
    for (int i0 = 0; i0 < n; i0 += nth) {
        const int cnt = std::min(nth, n - i0);
        std::vector<std::thread> th;
        for (int t = 0; t < cnt; ++t)
            th.emplace_back([&, t] { parts[t].n = 0; fmt(parts[t], i0 + t); });
        for (auto& x : th) x.join();
        for (int t = 0; t < cnt; ++t) fwrite(parts[t].b.data(), 1, parts[t].n, f);
    }
}

struct Shared {
    Options opt;
    int W = 0, H = 0, nCtus = 0, filterType = 0;
    const uint16_t* frames = nullptr;
    std::vector<std::vector<int32_t>> keepCost, keepSad, keepSatd;  // per frame, only those that get logged
    std::vector<std::vector<uint8_t>> keepMode;                     // per frame, with --DecisionsLog
    std::vector<std::vector<int32_t>> keepBest;
    std::atomic<int> errors{0};
    bool stamps = false;
    int binFd = -1;                                                 // --BinaryLog
};

constexpr int kBinHeader = 64;   // "MIPB200C", then u32 version, width, height, frames, CTUs, costs per CTU, bit depth, filter type, kernel index

// Engine of GPU g: context, streams, pinned rings, tables.  Runs before the timed window, like the reference's
// platform / queue / buffer / program setup (main.cpp:87-549).
mipb200_engine* create_engine(Shared* sh, int g, mipb200_config* cfg_out) {
    const Options& o = sh->opt;
    mipb200_config cfg;
    cfg.width = sh->W; cfg.height = sh->H; cfg.device = o.deviceIndex + g;
    cfg.filter_type = sh->filterType; cfg.kernel_idx = o.kernelIdx; cfg.slots = 3;
    cfg.top_k = o.topK > 1 ? o.topK : 0;
    cfg.bit_depth = o.bitDepth;
    const bool wantLog = !o.noLog, wantDec = !o.decisionsLog.empty(), wantBin = !o.binaryLog.empty();
    cfg.emit = (wantLog || wantBin || !wantDec ? MIPB200_EMIT_COSTS : 0) | (wantLog && !o.compat ? MIPB200_EMIT_SAD_SATD : 0) |
               (wantDec ? MIPB200_EMIT_DECISIONS : 0);
    mipb200_engine* e = nullptr;
    if (mipb200_create(&e, &cfg) != 0) {
        fprintf(stderr, "[!] ERROR (GPU %d): %s\n", cfg.device, mipb200_last_error());
        sh->errors++;
        return nullptr;
    }
    *cfg_out = cfg;
    return e;
}

// one host thread per GPU: frames poc = g, g+G, g+2G, ...
void gpu_worker(Shared* sh, mipb200_engine* e, mipb200_config cfg, int g, int G) {
    const Options& o = sh->opt;
    const bool wantLog = !o.noLog, wantDec = !o.decisionsLog.empty();
    const size_t fpx = (size_t)sh->W * sh->H, ncost = (size_t)sh->nCtus * MIP_COSTS_PER_CTU;
    int next = g, done = g;
    auto collect_one = [&]() -> bool {
        mipb200_result r;
        if (mipb200_collect(e, &r) != 0) { fprintf(stderr, "[!] ERROR (GPU %d): %s\n", cfg.device, mipb200_last_error()); sh->errors++; return false; }
        const int poc = (int)r.poc;
        if (sh->binFd >= 0) {   // raw table straight from the pinned ring to its place in the file (pwrite: any frame order, any thread)
            const char* src = reinterpret_cast<const char*>(r.cost);
            size_t left = ncost * sizeof(int32_t);
            off_t off = (off_t)kBinHeader + (off_t)poc * (off_t)left;
            while (left) {
                const ssize_t w = pwrite(sh->binFd, src, left, off);
                if (w <= 0) { perror("error while writing the binary log"); sh->errors++; return false; }
                src += w; off += w; left -= (size_t)w;
            }
        }
        if (wantLog && (poc == 0 || o.allFrames)) {   // the reference exports frame 0 only (main.cpp:1268)
            memcpy(sh->keepCost[poc].data(), r.cost, ncost * sizeof(int32_t));
            if (r.sad) { memcpy(sh->keepSad[poc].data(), r.sad, ncost * sizeof(int32_t)); memcpy(sh->keepSatd[poc].data(), r.satd, ncost * sizeof(int32_t)); }
        }
        if (wantDec) {
            const size_t ncu = (size_t)sh->nCtus * MIP_CUS_PER_CTU * (r.top_k ? r.top_k : 1);
            const uint8_t* bm = r.top_k ? r.topk_mode : r.best_mode;
            const int32_t* bc = r.top_k ? r.topk_cost : r.best_cost;
            memcpy(sh->keepMode[poc].data(), bm, ncu);
            memcpy(sh->keepBest[poc].data(), bc, ncu * sizeof(int32_t));
        }
        done += G;
        return true;
    };
    while (done < o.nFrames) {
        while (next < o.nFrames && mipb200_in_flight(e) < cfg.slots) {
            if (G == 1) printf("Current frame %d\n", next);
            if (sh->stamps) {
                // The reference stamps every stage's enqueue (main.cpp:738-1216); here one fused kernel is every stage, so
                // all stamps bracket the single asynchronous submit.  computeEnergy_NVIDIA.py:44-96 parses these names.
                if (next > 0) print_timestamp("START WRITE SAMPLES MEMOBJ");
                if (o.useAlt) print_timestamp("START ENQUEUE filterFrame");
                print_timestamp("START ENQUEUE initBoundaries");
                print_timestamp("START ENQUEUE reducedPred");
                print_timestamp("START ENQUEUE upsamplePred_SIZEID=2");
                print_timestamp("START ENQUEUE upsamplePred_SIZEID=1");
                print_timestamp("START ENQUEUE upsamplePred_SIZEID=0");
            }
            const int rcs = mipb200_submit(e, sh->frames + fpx * next, next);
            if (sh->stamps) {
                print_timestamp("FINISH WRITE SAMPLES MEMOBJ");
                if (o.useAlt) print_timestamp("FINISH ENQUEUE filterFrame");
                print_timestamp("FINISH ENQUEUE initBoundaries");
                print_timestamp("FINISH ENQUEUE reducedPred");
                print_timestamp("FINISH ENQUEUE upsamplePred_SIZEID=2");
                print_timestamp("FINISH ENQUEUE upsamplePred_SIZEID=1");
                print_timestamp("FINISH ENQUEUE upsamplePred_SIZEID=0");
            }
            if (rcs != 0) {
                fprintf(stderr, "[!] ERROR (GPU %d): %s\n", cfg.device, mipb200_last_error());
                sh->errors++;
                return;
            }
            next += G;
        }
        if (sh->stamps) print_timestamp("START READ DISTORTION");
        if (!collect_one()) break;
        if (sh->stamps && done < o.nFrames) print_timestamp("FINISH READ DISTORTION");
    }
}

}  // namespace

int main(int argc, char** argv) {
    Shared sh;
    Options& o = sh.opt;
    if (!parse_args(argc, argv, o)) return 1;
    if (o.help) { print_help(); return 1; }

    int po_error = report_parameters(o);
    if (o.useAlt) {
        int ft = 0;
        for (int i = 0; i < 8; ++i)
            if (o.filter == MIP_FILTER_NAMES[i]) ft = i + 1;
        if (!ft) {   // main.cpp:74-77
            printf("  [!] ERROR: Filter type %s not supported\n", o.filter.c_str());
            return 0;
        }
        sh.filterType = ft;
    }
    if (po_error > 0) {
        printf("Exiting after finding errors in input parameters\n");
        return 1;
    }
    print_timestamp("STARTED HOST");

    int W = 0, H = 0;
    {
        size_t x = o.resolution.find('x');
        if (x == std::string::npos || !to_int(o.resolution.substr(0, x), &W) || !to_int(o.resolution.substr(x + 1), &H)) {
            printf("  [!] ERROR: Input resolution \"%s\" not set properly\n", o.resolution.c_str());
            return 0;
        }
    }
    if (W < 8 || H < 4 || W % 8 != 0 || H % 4 != 0) {
        printf("[!] ERROR: Unsupported resolution %dx%d\n", W, H);
        printf("Supported resolutions are: any WxH with W %% 8 == 0 and H %% 4 == 0, e.g.\n  3840x2160\n  1920x1080\n  1280x720\n  832x480\n  416x240\n");
        return 0;
    }
    if (o.nFrames < 1 || o.numGpus < 1) { printf("  [!] ERROR: FramesToBeEncoded and NumGpus must be positive\n"); return 1; }
    if (o.topK < 0 || o.topK > MIPB200_TOPK_MAX) { printf("  [!] ERROR: TopK must be in 1..%d\n", MIPB200_TOPK_MAX); return 1; }
    if (o.bitDepth != 8 && o.bitDepth != 10 && o.bitDepth != 12) { printf("  [!] ERROR: BitDepth must be 8, 10 or 12\n"); return 1; }
    if (o.topK > 1 && o.decisionsLog.empty()) { printf("  [!] ERROR: TopK needs --DecisionsLog\n"); return 1; }
    sh.stamps = o.stageStamps < 0 ? (TRACE_POWER && o.numGpus == 1) : o.stageStamps != 0;
    sh.W = W; sh.H = H; sh.nCtus = mipb200_num_ctus(W, H);

    print_timestamp("START READ SAMPLES .csv");
    std::vector<uint16_t> frames;
    if (o.inputFormat == "csv") {
        if (!read_frames_csv(o.input, W, H, o.nFrames, frames)) return 1;
    } else if (o.inputFormat == "u16" || o.inputFormat == "yuv420p" || o.inputFormat == "yuv420p10le") {
        if (!read_frames_binary(o.input, o.inputFormat, W, H, o.nFrames, frames)) return 1;
    } else {
        printf("  [!] ERROR: InputFormat %s not supported (csv, u16, yuv420p, yuv420p10le)\n", o.inputFormat.c_str());
        return 1;
    }
    print_timestamp("FINISH READ SAMPLES .csv");
    sh.frames = frames.data();
    sh.keepCost.resize(o.nFrames); sh.keepSad.resize(o.nFrames); sh.keepSatd.resize(o.nFrames);
    sh.keepMode.resize(o.nFrames); sh.keepBest.resize(o.nFrames);

    // ---- timed window: first upload -> last result resident on the host (main.cpp:566-569, 1247-1250)
    {   // device selection banner of the reference (main.cpp:220-228); it exits with 0 on a bad index
        const int found = mipb200_device_count();
        if (found < 0) { fprintf(stderr, "[!] ERROR: %s\n", mipb200_last_error()); return 1; }
        if (o.deviceIndex < 0 || o.deviceIndex + o.numGpus > found) {
            printf("Incorrect GPU index. Only %d GPUs are detected\n", found);
            return 0;
        }
        for (int g = 0; g < o.numGpus; ++g) printf("COMPUTING ON GPU %d\n", o.deviceIndex + g);
    }
    // ---- set-up (not timed, like the reference's context / buffer / program creation)
    print_timestamp("START BUILD KERNELS");
    std::vector<mipb200_engine*> engines(o.numGpus, nullptr);
    std::vector<mipb200_config> cfgs(o.numGpus);
    {
        std::vector<std::thread> th;
        for (int g = 0; g < o.numGpus; ++g) th.emplace_back([&, g] { engines[g] = create_engine(&sh, g, &cfgs[g]); });
        for (auto& t : th) t.join();
    }
    auto destroy_all = [&] { for (auto* e : engines) mipb200_destroy(e); };
    if (sh.errors) { destroy_all(); return 1; }
    // host result arrays, allocated and touched here like the reference's return_* arrays (main.cpp:656-662)
    {
        const size_t ncost = (size_t)sh.nCtus * MIP_COSTS_PER_CTU, ncu = (size_t)sh.nCtus * MIP_CUS_PER_CTU * (o.topK > 1 ? o.topK : 1);
        for (int poc = 0; poc < o.nFrames; ++poc) {
            if (!o.noLog && (poc == 0 || o.allFrames)) {
                sh.keepCost[poc].resize(ncost);
                if (!o.compat) { sh.keepSad[poc].resize(ncost); sh.keepSatd[poc].resize(ncost); }
            }
            if (!o.decisionsLog.empty()) { sh.keepMode[poc].resize(ncu); sh.keepBest[poc].resize(ncu); }
        }
    }
    if (!o.binaryLog.empty()) {
        sh.binFd = open(o.binaryLog.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
        if (sh.binFd < 0) { perror("error while opening the binary log"); destroy_all(); return 1; }
        uint32_t hdr[kBinHeader / 4] = {0};
        memcpy(hdr, "MIPB200C", 8);
        hdr[2] = 1; hdr[3] = (uint32_t)W; hdr[4] = (uint32_t)H; hdr[5] = (uint32_t)o.nFrames; hdr[6] = (uint32_t)sh.nCtus;
        hdr[7] = MIP_COSTS_PER_CTU; hdr[8] = (uint32_t)o.bitDepth; hdr[9] = (uint32_t)sh.filterType; hdr[10] = (uint32_t)o.kernelIdx;
        if (pwrite(sh.binFd, hdr, sizeof(hdr), 0) != (ssize_t)sizeof(hdr)) { perror("error while writing the binary log"); destroy_all(); return 1; }
    }
    // page-lock the frames so that every upload is a DMA from where the samples already are (no staging copy)
    const bool pinned = mipb200_pin_host(frames.data(), frames.size() * sizeof(uint16_t)) == 0;
    print_timestamp("FINISH BUILD KERNELS");

    std::vector<unsigned long long> mj0(o.numGpus, 0), mj1(o.numGpus, 0);
    bool haveEnergy = o.energy;
    for (int g = 0; g < o.numGpus && haveEnergy; ++g)
        if (mipb200_device_energy_mj(o.deviceIndex + g, &mj0[g]) != 0) {
            printf("  [!] Energy counter unavailable: %s\n", mipb200_last_error());
            haveEnergy = false;
        }
    print_timestamp("START WRITE SAMPLES MEMOBJ");
    const double t0 = now_ms();
    {
        std::vector<std::thread> th;
        for (int g = 0; g < o.numGpus; ++g) th.emplace_back(gpu_worker, &sh, engines[g], cfgs[g], g, o.numGpus);
        for (auto& t : th) t.join();
    }
    const double t1 = now_ms();
    print_timestamp("FINISH READ DISTORTION");
    for (int g = 0; g < o.numGpus && haveEnergy; ++g)
        if (mipb200_device_energy_mj(o.deviceIndex + g, &mj1[g]) != 0) haveEnergy = false;
    destroy_all();
    if (sh.binFd >= 0) close(sh.binFd);
    if (pinned) mipb200_unpin_host(frames.data());
    if (sh.errors) return 1;

    if (!o.noLog) {
        // like the reference, a log is written even when -l is omitted (file ".csv", main.cpp:1264-1269)
        std::string name = o.prefix + ".csv";
        FILE* f = fopen(name.c_str(), "w");
        if (!f) { perror("error while opening the output file"); return 1; }
        LogBuf lb(f);
        const char* hdr = o.allFrames ? "POC,CTU,cuSizeName,W,H,CU,X,Y,Mode,SAD,SATD,minSadHad\n" : "CTU,cuSizeName,W,H,CU,X,Y,Mode,SAD,SATD,minSadHad\n";
        lb.put(hdr, strlen(hdr));
        lb.flush();
        // 13.2 M lines per 1080p frame: CTUs are formatted by all host threads into private buffers (one CTU = 4.4 MB of
        // text each) and written in CTU order
        for (int poc = 0; poc < (o.allFrames ? o.nFrames : 1); ++poc) {
            const int32_t* cst = sh.keepCost[poc].data();
            const int32_t* sd = sh.keepSad[poc].empty() ? nullptr : sh.keepSad[poc].data();
            const int32_t* st = sh.keepSatd[poc].empty() ? nullptr : sh.keepSatd[poc].data();
            format_parallel(f, sh.nCtus, 6u << 20, [&](LogBuf& b, int ctu) { write_frame_log(b, poc, o.allFrames, cst, sd, st, ctu, ctu + 1, W, o.compat); });
        }
        fclose(f);
    }

    if (!o.decisionsLog.empty()) {
        FILE* f = fopen(o.decisionsLog.c_str(), "w");
        if (!f) { perror("error while opening the decisions log"); return 1; }
        LogBuf lb(f);
        const int k = o.topK > 1 ? o.topK : 1;
        std::string hdr = "POC,CTU,cuSizeName,W,H,CU,X,Y,BestMode,BestCost";
        for (int j = 2; j <= k; ++j) hdr += ",Mode" + std::to_string(j) + ",Cost" + std::to_string(j);
        hdr += "\n";
        lb.put(hdr.c_str(), hdr.size());
        lb.flush();
        format_parallel(f, o.nFrames, 8u << 20, [&](LogBuf& b, int poc) { write_decisions(b, poc, sh.keepMode[poc].data(), sh.keepBest[poc].data(), k, 0, sh.nCtus, W); });
        fclose(f);
    }

    // reportTimingResults_Compact (main_aux_functions.h:908-914)
    printf("=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=\n");
    printf("TIMING RESULTS (miliseconds)\n");
    printf("Elapsed time (ms) from writing samples to reading distortion (%dx), %d\n", o.nFrames, (int)lround(t1 - t0));
    printf("=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=\n\n");
    printf("Throughput: %.1f frames/s on %d GPU(s)\n", o.nFrames * 1e3 / (t1 - t0 > 0 ? t1 - t0 : 1e-3), o.numGpus);
    if (o.energy && haveEnergy) {
        // the board's own energy counter over the timed window (the reference integrates an nvidia-smi power trace
        // between the same two stamps, computeEnergy_NVIDIA.py:98-140)
        double joules = 0;
        for (int g = 0; g < o.numGpus; ++g) joules += (double)(mj1[g] - mj0[g]) * 1e-3;
        printf("ENERGY RESULTS\n");
        printf("Energy (J) from writing samples to reading distortion (%dx), %.3f\n", o.nFrames, joules);
        printf("Energy per frame (J), %.4f\n", joules / o.nFrames);
        printf("Average power (W), %.1f\n", joules / ((t1 - t0 > 0 ? t1 - t0 : 1e-3) * 1e-3));
    }
    print_timestamp("FINISHED HOST");
    return 0;
}
