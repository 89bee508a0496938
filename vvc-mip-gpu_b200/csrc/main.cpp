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
//
// Memory model.  The reference reads every frame into one array and keeps every frame's table (main.cpp:364-384,
// :656-662); that is 250 GB for BASELINE config 5 (2048 frames of 7680x4320).  Here frames live in a bounded,
// page-locked RING (--RingFrames, default: 2 GiB worth, at least Slots + 1 per GPU) that the engines DMA from in place.  When
// the input fits the ring it is loaded before the timed window, exactly like the reference; otherwise a reader thread
// streams it (the ring is full when the window opens, slots are refilled as their frames are collected).  Results are
// consumed frame by frame as they are collected -- written to their place in the binary logs (pwrite), hashed into the
// digest, or formatted and appended in POC order to the text logs -- so host memory does not grow with -f.
//
// Extensions (all off by default): --NumGpus G (frames sharded poc % G over G GPUs, one host thread each, no
// collective), --AllFrames (log every frame with a leading POC column), --Compat (print 0 in the SAD/SATD columns like
// the reference's MAX_PERFORMANCE_DIST build), --NoLog (skip the text log), --InputFormat csv|u16|yuv420p|yuv420p10le
// (binary luma input instead of the 2 M stoi() calls per 1080p frame), --InputFrames P (the file holds P frames, frame
// poc is file frame poc % P: a pool cycled like SURVEY 8(d) configs 4/5), --DecisionsLog FILE (text: per-CU best mode
// + cost of EVERY frame keyed by POC,X,Y,W,H), --DecisionsBin FILE (the same table raw: 5 bytes per CU), --TopK k
// (shortlists in both), --BinaryLog FILE (raw int32 cost tables of every frame), --Digest FILE (one 64-bit hash per
// frame and result array: what G-independence is checked with), --BitDepth 8|10|12, --Energy (NVML joules per
// frame), --StageStamps 0|1 (the reference's TRACE_POWER stamps).
//
// The device work goes through the C ABI of include/mipb200.h only.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mipb200.h"
#include "mip_compact.h"
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
    int inputFrames = 0;    // frames the input file holds (0: as many as -f); frame poc is file frame poc % inputFrames
    int ringFrames = 0;     // capacity of the page-locked frame ring (0: 2 GiB worth, at least Slots + 1 per GPU)
    int slots = 3;          // frames in flight per engine
    std::string inputFormat = "csv", decisionsLog, binaryLog, decisionsBin, digest, compactLog;
    bool allFrames = false, compat = false, noLog = false, help = false, energy = false;
};

const char* kLongOpts[] = {"help", "DeviceIndex", "FramesToBeEncoded", "Resolution", "OriginalFrames", "OutputPreffix",
                           "FilterType", "KernelIdx", "UseAlternativeSamples", "NumGpus", "AllFrames", "Compat", "NoLog",
                           "InputFormat", "DecisionsLog", "TopK", "Energy", "StageStamps", "BitDepth", "BinaryLog", "DecisionsBin", "Digest",
                           "InputFrames", "RingFrames", "Slots", "CompactLog"};
const bool kTakesValue[] = {false, true, true, true, true, true, true, true, true, true, false, false, false, true, true,
                            true, false, true, true, true, true, true, true, true, true, true};
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
           "  --CompactLog arg               like --BinaryLog with the compact table (uint16 for CUs of <= 32 samples): 71 %% of the bytes\n"
           "  --DecisionsBin arg             write every frame's decisions raw: uint8 modes then int32 costs (64-byte header)\n"
           "  --Digest arg                   write POC,<64-bit hash of each result array> for every frame\n"
           "  --InputFrames arg              frames held by the input file; frame poc reads file frame poc %% arg (default: -f)\n"
           "  --RingFrames arg               capacity of the page-locked frame ring (default: 2 GiB worth, >= Slots + 1 per GPU)\n"
           "  --Slots arg (=3)               frames in flight per GPU (upload, kernels and read-back of different frames overlap)\n"
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
            case 20: o.decisionsBin = val; break;
            case 21: o.digest = val; break;
            case 22: ok = to_int(val, &o.inputFrames); break;
            case 23: ok = to_int(val, &o.ringFrames); break;
            case 24: ok = to_int(val, &o.slots); break;
            case 25: o.compactLog = val; break;
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

// Sequential reader of luma frames, one frame per call (nothing but the current frame is held):
//   csv          N*H lines of W comma-separated integers, extra fields ignored (main.cpp:364-384)
//   u16          W*H little-endian uint16 per frame
//   yuv420p      8-bit planar 4:2:0, chroma skipped        yuv420p10le  16-bit planar 4:2:0, chroma skipped
// Sample values are taken as they are (the pipeline is 10-bit by default, intra.cl:61) but must fit --BitDepth: larger
// values would overflow the engine's packed 16-bit arithmetic silently, so they are an input error here.
class FrameSource {
  public:
    ~FrameSource() { if (f_) fclose(f_); }
    bool open(const std::string& path, const std::string& fmt, int W, int H, int bitDepth, long framesWanted) {
        path_ = path; fmt_ = fmt; W_ = W; H_ = H; bits_ = bitDepth; wanted_ = framesWanted;
        f_ = fopen(path.c_str(), "rb");
        if (!f_) { perror("error while opening samples files"); return false; }
        if (fmt == "csv") buf_.resize(4u << 20);
        else if (fmt == "yuv420p") tmp_.resize((size_t)W * H);
        return true;
    }
    void rewind_to_start() { fseek(f_, 0, SEEK_SET); beg_ = end_ = 0; line_ = 0; frame_ = 0; }
    // next frame of the file into dst[H][W]; false (message on stderr) on a short file, a malformed field or an out-of-range sample
    bool next(uint16_t* dst) {
        const bool ok = fmt_ == "csv" ? next_csv(dst) : next_binary(dst);
        if (!ok) return false;
        unsigned all = 0;
        const size_t px = (size_t)W_ * H_;
        for (size_t i = 0; i < px; ++i) all |= dst[i];
        if (all >> bits_) {
            size_t i = 0;
            while (!(dst[i] >> bits_)) ++i;
            fprintf(stderr, "[!] ERROR: frame %ld of %s: sample %u at row %zu, column %zu does not fit %d bits (see --BitDepth)\n", frame_, path_.c_str(),
                    (unsigned)dst[i], i / W_, i % W_, bits_);
            return false;
        }
        ++frame_;
        return true;
    }

  private:
    int getc_buf() {
        if (beg_ == end_) {
            end_ = fread(buf_.data(), 1, buf_.size(), f_);
            beg_ = 0;
            if (end_ == 0) return -1;
        }
        return (unsigned char)buf_[beg_++];
    }
    bool next_csv(uint16_t* dst) {
        for (int y = 0; y < H_; ++y, ++line_) {
            int c = getc_buf();
            if (c < 0) {
                fprintf(stderr, "[!] ERROR: %s holds %ld lines, need %ld (%ld frames of %d rows)\n", path_.c_str(), line_, wanted_ * H_, wanted_, H_);
                return false;
            }
            uint16_t* row = dst + (size_t)y * W_;
            for (int x = 0; x < W_; ++x) {
                while (c == ' ' || c == '\t') c = getc_buf();
                if (c < '0' || c > '9') {
                    fprintf(stderr, "[!] ERROR: line %ld of %s: field %d is not a number (need %d samples per line)\n", line_ + 1, path_.c_str(), x + 1, W_);
                    return false;
                }
                unsigned v = 0;
                while (c >= '0' && c <= '9') { v = v * 10 + (unsigned)(c - '0'); if (v > 65535u) v = 65535u; c = getc_buf(); }
                row[x] = (uint16_t)v;
                if (c == ',') c = getc_buf();
            }
            while (c >= 0 && c != '\n') c = getc_buf();
        }
        return true;
    }
    bool next_binary(uint16_t* dst) {
        const size_t px = (size_t)W_ * H_;
        const bool eight = fmt_ == "yuv420p";
        const size_t chroma = fmt_ == "u16" ? 0 : px / 2 * (eight ? 1 : 2);
        size_t got;
        if (eight) {
            got = fread(tmp_.data(), 1, px, f_);
            for (size_t i = 0; i < px; ++i) dst[i] = tmp_[i];
        } else {
            got = fread(dst, 2, px, f_);
        }
        if (got != px) {
            fprintf(stderr, "[!] ERROR: %s holds fewer than %ld frames of %dx%d (%s)\n", path_.c_str(), wanted_, W_, H_, fmt_.c_str());
            return false;
        }
        return !chroma || fseek(f_, (long)chroma, SEEK_CUR) == 0;
    }
    FILE* f_ = nullptr;
    std::string path_, fmt_;
    int W_ = 0, H_ = 0, bits_ = 10;
    long wanted_ = 0, line_ = 0, frame_ = 0;
    std::vector<char> buf_;
    size_t beg_ = 0, end_ = 0;
    std::vector<uint8_t> tmp_;
};

// Page-locked ring the engines DMA frames from in place.  `period` = distinct frames (-f, or --InputFrames when smaller).
//  * resident (period <= capacity): slot = poc % period, filled once before the timed window, never released;
//  * streamed: slot = poc % capacity; the reader thread fills frames in POC order as slots are released by the worker
//    that collected their frame.  Capacity >= 4 frames per GPU keeps the reader ahead of 3 frames in flight per engine.
struct FrameRing {
    int capacity = 0, period = 0;
    bool resident = true;
    size_t fpx = 0;
    uint16_t* base = nullptr;
    bool pinned = false;
    std::vector<long> holds;      // POC whose samples the slot holds (-1: none yet)
    std::vector<char> busy;       // streamed mode: filled and not yet released
    long filled = 0;              // frames read so far
    bool failed = false;
    std::mutex mu;
    std::condition_variable cv;

    int slot_of(long poc) const { return (int)(resident ? poc % period : poc % capacity); }
    uint16_t* slot_ptr(int s) const { return base + fpx * (size_t)s; }
    // worker: wait until frame poc is in the ring; nullptr if the reader failed
    const uint16_t* acquire(long poc) {
        const int s = slot_of(poc);
        if (resident) return slot_ptr(s);
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return failed || (holds[s] == poc && busy[s]); });
        return failed ? nullptr : slot_ptr(s);
    }
    void release(long poc) {
        if (resident) return;
        std::lock_guard<std::mutex> lk(mu);
        busy[slot_of(poc)] = 0;
        cv.notify_all();
    }
    void fail() { std::lock_guard<std::mutex> lk(mu); failed = true; cv.notify_all(); }
};

// 64-bit digest of a result array, sensitive to any changed, moved or missing word.  Eight accumulator lanes over 64-byte
// stripes (acc[i] += lo32(w ^ k) * hi32(w ^ k); acc[i ^ 1] += w: compiles to packed 32x32->64 multiplies), scrambled every
// 64 KiB.  Arrays are hashed in fixed 4 MiB segments and the digest is the digest of the segment digests, so large arrays
// (55 MB of decisions per 4320p frame, 7.7 GB/s per GPU) can be hashed by a few threads with the same result as by one.
uint64_t digest64_segment(const void* data, size_t bytes) {
    static const uint64_t K[8] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0x27D4EB2F165667C5ull,
                                  0xFF51AFD7ED558CCDull, 0xC4CEB9FE1A85EC53ull, 0x2545F4914F6CDD1Dull, 0x94D049BB133111EBull};
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint64_t acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = K[7 - i];
    const size_t stripes = bytes / 64;
    for (size_t s = 0; s < stripes; ++s, p += 64) {
        uint64_t w[8];
        memcpy(w, p, 64);
        for (int i = 0; i < 8; ++i) {
            const uint64_t d = w[i] ^ K[i];
            acc[i] += (d & 0xffffffffu) * (d >> 32);
            acc[i ^ 1] += w[i];
        }
        if ((s & 1023) == 1023)
            for (int i = 0; i < 8; ++i) acc[i] = (acc[i] ^ (acc[i] >> 47) ^ K[i]) * 0x9E3779B1ull;
    }
    uint64_t tail[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    memcpy(tail, p, bytes % 64);
    for (int i = 0; i < 8; ++i) { const uint64_t d = tail[i] ^ K[i]; acc[i] += (d & 0xffffffffu) * (d >> 32); acc[i ^ 1] += tail[i]; }
    uint64_t r = bytes * 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 8; ++i) { r = (r ^ acc[i]) * 0xFF51AFD7ED558CCDull; r ^= r >> 29; }
    return r;
}

uint64_t digest64(const void* data, size_t bytes, int maxThreads = 4) {
    constexpr size_t SEG = (size_t)4 << 20;
    const size_t nseg = (bytes + SEG - 1) / SEG;
    std::vector<uint64_t> part(nseg ? nseg : 1, 0);
    const uint8_t* p = static_cast<const uint8_t*>(data);
    auto seg = [&](size_t i) { part[i] = digest64_segment(p + i * SEG, std::min(SEG, bytes - i * SEG)); };
    const int nth = (int)std::min<size_t>(maxThreads, nseg);
    if (nth <= 1) {
        for (size_t i = 0; i < nseg; ++i) seg(i);
    } else {
        std::atomic<size_t> next{0};
        auto work = [&] { for (size_t i; (i = next.fetch_add(1)) < nseg;) seg(i); };
        std::vector<std::thread> th;
        for (int t = 1; t < nth; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    return digest64_segment(part.data(), part.size() * sizeof(uint64_t)) ^ (uint64_t)bytes;
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

// A text log that frames reach out of order (one worker per GPU) but that must be written in POC order: a worker waits
// for its frame's turn, formats and writes (in bounded rounds, all host threads formatting), and passes the turn on.
// Workers take their frames in increasing POC order and the frame ring holds more frames than can be in flight, so the
// smallest outstanding POC can always proceed: no deadlock.
struct OrderedLog {
    FILE* f = nullptr;
    long next = 0;
    bool failed = false;
    std::mutex mu;
    std::condition_variable cv;
    bool begin(long poc) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return failed || next == poc; });
        return !failed;
    }
    void end() { std::lock_guard<std::mutex> lk(mu); ++next; cv.notify_all(); }
    void abort() { std::lock_guard<std::mutex> lk(mu); failed = true; cv.notify_all(); }
};

struct Shared {
    Options opt;
    int W = 0, H = 0, nCtus = 0, filterType = 0, k = 1;             // k: entries per CU in the decision arrays (--TopK)
    unsigned emit = 0;                                              // MIPB200_EMIT_* of every engine
    FrameRing ring;
    std::vector<int32_t> keepCost, keepSad, keepSatd;               // frame 0 only: the reference's log (main.cpp:1268)
    std::vector<uint64_t> digCost, digMode, digBest;                // --Digest: one hash per frame and array
    std::atomic<int> errors{0};
    bool stamps = false;
    int binFd = -1, decFd = -1, cmpFd = -1;                         // --BinaryLog, --DecisionsBin, --CompactLog
    OrderedLog costLog, decLog;                                     // --AllFrames text log, --DecisionsLog
};

constexpr int kBinHeader = 64;   // "MIPB200C", then u32 version, width, height, frames, CTUs, costs per CTU, bit depth, filter type, kernel index
                                 // "MIPB200D", then u32 version, width, height, frames, CTUs, CUs per CTU, bit depth, filter type, kernel index, k

bool pwrite_all(int fd, const void* data, size_t bytes, off_t off) {
    const char* src = static_cast<const char*>(data);
    while (bytes) {
        const ssize_t w = pwrite(fd, src, bytes, off);
        if (w <= 0) return false;
        src += w; off += w; bytes -= (size_t)w;
    }
    return true;
}

// Engine of GPU g: context, streams, pinned rings, tables.  Runs before the timed window, like the reference's
// platform / queue / buffer / program setup (main.cpp:87-549).
mipb200_engine* create_engine(Shared* sh, int g, mipb200_config* cfg_out) {
    const Options& o = sh->opt;
    mipb200_config cfg;
    cfg.width = sh->W; cfg.height = sh->H; cfg.device = o.deviceIndex + g;
    cfg.filter_type = sh->filterType; cfg.kernel_idx = o.kernelIdx; cfg.slots = o.slots;
    cfg.top_k = o.topK > 1 ? o.topK : 0;
    cfg.bit_depth = o.bitDepth;
    cfg.emit = sh->emit;
    mipb200_engine* e = nullptr;
    if (mipb200_create(&e, &cfg) != 0) {
        fprintf(stderr, "[!] ERROR (GPU %d): %s\n", cfg.device, mipb200_last_error());
        sh->errors++;
        return nullptr;
    }
    *cfg_out = cfg;
    return e;
}

// What happens to a frame's results, on the worker thread that collected them, while the GPU works on the next frames.
bool consume_result(Shared* sh, const mipb200_result& r) {
    const Options& o = sh->opt;
    const long poc = (long)r.poc;
    const size_t ncost = (size_t)sh->nCtus * MIP_COSTS_PER_CTU, ncu = (size_t)sh->nCtus * MIP_CUS_PER_CTU * sh->k;
    const uint8_t* bm = r.top_k ? r.topk_mode : r.best_mode;
    const int32_t* bc = r.top_k ? r.topk_cost : r.best_cost;
    if (sh->binFd >= 0 && !pwrite_all(sh->binFd, r.cost, ncost * sizeof(int32_t), (off_t)kBinHeader + (off_t)poc * (off_t)(ncost * sizeof(int32_t)))) {
        perror("error while writing the binary log");   // raw table straight from the pinned ring to its place in the file
        return false;
    }
    if (sh->cmpFd >= 0) {
        const size_t rec = (size_t)sh->nCtus * MIP_COMPACT_BYTES_PER_CTU;
        if (!pwrite_all(sh->cmpFd, r.cost_compact, rec, (off_t)kBinHeader + (off_t)poc * (off_t)rec)) { perror("error while writing the compact log"); return false; }
    }
    if (sh->decFd >= 0) {   // frame record: modes [nCTU][5380][k] uint8, then costs [nCTU][5380][k] int32
        const off_t rec = (off_t)(ncu * 5), off = (off_t)kBinHeader + (off_t)poc * rec;
        if (!pwrite_all(sh->decFd, bm, ncu, off) || !pwrite_all(sh->decFd, bc, ncu * sizeof(int32_t), off + (off_t)ncu)) {
            perror("error while writing the binary decisions");
            return false;
        }
    }
    if (!o.digest.empty()) {
        if (r.cost) sh->digCost[poc] = digest64(r.cost, ncost * sizeof(int32_t));
        if (r.cost_compact) sh->digCost[poc] = digest64(r.cost_compact, (size_t)sh->nCtus * MIP_COMPACT_BYTES_PER_CTU);
        if (bm) { sh->digMode[poc] = digest64(bm, ncu); sh->digBest[poc] = digest64(bc, ncu * sizeof(int32_t)); }
    }
    if (!o.noLog && !o.allFrames && poc == 0) {   // the reference exports frame 0 only (main.cpp:1268), after the timed window
        memcpy(sh->keepCost.data(), r.cost, ncost * sizeof(int32_t));
        if (r.sad) { memcpy(sh->keepSad.data(), r.sad, ncost * sizeof(int32_t)); memcpy(sh->keepSatd.data(), r.satd, ncost * sizeof(int32_t)); }
    }
    if (!o.noLog && o.allFrames) {   // 13.2 M lines per 1080p frame: one CTU (4.4 MB of text) per item, appended in POC order
        if (!sh->costLog.begin(poc)) return false;
        format_parallel(sh->costLog.f, sh->nCtus, 6u << 20, [&](LogBuf& b, int ctu) { write_frame_log(b, poc, true, r.cost, r.sad, r.satd, ctu, ctu + 1, sh->W, o.compat); });
        sh->costLog.end();
    }
    if (!o.decisionsLog.empty()) {   // 726 300 lines per 1080p frame: 16 CTUs per item
        if (!sh->decLog.begin(poc)) return false;
        const int per = 16, items = (sh->nCtus + per - 1) / per;
        format_parallel(sh->decLog.f, items, 8u << 20, [&](LogBuf& b, int it) { write_decisions(b, poc, bm, bc, sh->k, it * per, std::min(sh->nCtus, (it + 1) * per), sh->W); });
        sh->decLog.end();
    }
    return true;
}

// one host thread per GPU: frames poc = g, g+G, g+2G, ...
void gpu_worker(Shared* sh, mipb200_engine* e, mipb200_config cfg, int g, int G) {
    const Options& o = sh->opt;
    long next = g, done = g;
    auto fail = [&](const char* what) {
        if (what) fprintf(stderr, "[!] ERROR (GPU %d): %s\n", cfg.device, what);
        sh->errors++;
        sh->ring.fail();
        sh->costLog.abort();
        sh->decLog.abort();
    };
    while (done < o.nFrames) {
        while (next < o.nFrames && mipb200_in_flight(e) < cfg.slots) {
            const uint16_t* frame = sh->ring.acquire(next);
            if (!frame) { sh->errors++; return; }   // the reader failed (it said why)
            if (G == 1) printf("Current frame %ld\n", next);
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
            const int rcs = mipb200_submit(e, frame, next);
            if (sh->stamps) {
                print_timestamp("FINISH WRITE SAMPLES MEMOBJ");
                if (o.useAlt) print_timestamp("FINISH ENQUEUE filterFrame");
                print_timestamp("FINISH ENQUEUE initBoundaries");
                print_timestamp("FINISH ENQUEUE reducedPred");
                print_timestamp("FINISH ENQUEUE upsamplePred_SIZEID=2");
                print_timestamp("FINISH ENQUEUE upsamplePred_SIZEID=1");
                print_timestamp("FINISH ENQUEUE upsamplePred_SIZEID=0");
            }
            if (rcs != 0) { fail(mipb200_last_error()); return; }
            next += G;
        }
        if (sh->stamps) print_timestamp("START READ DISTORTION");
        mipb200_result r;
        if (mipb200_collect(e, &r) != 0) { fail(mipb200_last_error()); return; }
        sh->ring.release((long)r.poc);        // the frame has been uploaded and used: its ring slot may be refilled
        if (!consume_result(sh, r)) { fail(nullptr); return; }
        done += G;
        if (sh->stamps && done < o.nFrames) print_timestamp("FINISH READ DISTORTION");
    }
}

// streamed ring: frames in POC order into slot poc % capacity as slots are released
void reader_thread(Shared* sh, FrameSource* src) {
    FrameRing& ring = sh->ring;
    const Options& o = sh->opt;
    for (long poc = 0; poc < o.nFrames; ++poc) {
        const int s = ring.slot_of(poc);
        {
            std::unique_lock<std::mutex> lk(ring.mu);
            ring.cv.wait(lk, [&] { return ring.failed || !ring.busy[s]; });
            if (ring.failed) return;
        }
        if (poc > 0 && poc % ring.period == 0) src->rewind_to_start();
        if (!src->next(ring.slot_ptr(s))) { sh->errors++; ring.fail(); return; }
        std::lock_guard<std::mutex> lk(ring.mu);
        ring.holds[s] = poc;
        ring.busy[s] = 1;
        ring.filled = poc + 1;
        ring.cv.notify_all();
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
    if (o.topK > 1 && o.decisionsLog.empty() && o.decisionsBin.empty()) { printf("  [!] ERROR: TopK needs --DecisionsLog\n"); return 1; }
    if (o.slots < 1 || o.slots > 16) { printf("  [!] ERROR: Slots must be in 1..16\n"); return 1; }
    if (o.inputFrames < 0 || o.ringFrames < 0) { printf("  [!] ERROR: InputFrames and RingFrames must be positive\n"); return 1; }
    if (o.inputFormat != "csv" && o.inputFormat != "u16" && o.inputFormat != "yuv420p" && o.inputFormat != "yuv420p10le") {
        printf("  [!] ERROR: InputFormat %s not supported (csv, u16, yuv420p, yuv420p10le)\n", o.inputFormat.c_str());
        return 1;
    }
    sh.stamps = o.stageStamps < 0 ? (TRACE_POWER && o.numGpus == 1) : o.stageStamps != 0;
    sh.W = W; sh.H = H; sh.nCtus = mipb200_num_ctus(W, H);
    sh.k = o.topK > 1 ? o.topK : 1;
    {
        const bool wantLog = !o.noLog, wantDec = !o.decisionsLog.empty() || !o.decisionsBin.empty(), wantBin = !o.binaryLog.empty();
        const bool wantCmp = !o.compactLog.empty();
        if (wantCmp && (wantLog || wantBin || o.bitDepth == 12 || o.topK > 1)) {
            printf("  [!] ERROR: CompactLog needs --NoLog and excludes --BinaryLog, --TopK and --BitDepth=12 (it replaces the int32 table)\n");
            return 1;
        }
        const bool wantCost = wantLog || wantBin || (!wantDec && !wantCmp && o.digest.empty());    // nothing else asked for: the reference's table
        sh.emit = (wantCost ? MIPB200_EMIT_COSTS : 0) | (wantCmp ? MIPB200_EMIT_COSTS_COMPACT : 0) | (wantLog && !o.compat ? MIPB200_EMIT_SAD_SATD : 0) |
                  (wantDec || (!wantCost && !wantCmp) || !o.digest.empty() ? MIPB200_EMIT_DECISIONS : 0);
    }

    // ---- the frame ring: resident when the distinct frames fit, streamed otherwise
    FrameRing& ring = sh.ring;
    ring.fpx = (size_t)W * H;
    ring.period = o.inputFrames > 0 ? std::min(o.inputFrames, o.nFrames) : o.nFrames;
    {
        const size_t frameBytes = ring.fpx * sizeof(uint16_t);
        const long perGpu = o.slots + 1;   // frames in flight per engine plus one being read: the reader stays ahead
        long cap = o.ringFrames > 0 ? o.ringFrames : (long)std::max<size_t>((size_t)perGpu * o.numGpus, ((size_t)2 << 30) / frameBytes);
        cap = std::max<long>(cap, perGpu * o.numGpus);
        ring.resident = ring.period <= cap;
        ring.capacity = (int)(ring.resident ? ring.period : cap);
        void* mem = nullptr;
        if (posix_memalign(&mem, 4096, frameBytes * (size_t)ring.capacity) != 0) {
            printf("  [!] ERROR: cannot allocate the frame ring (%d frames of %zu bytes)\n", ring.capacity, frameBytes);
            return 1;
        }
        ring.base = static_cast<uint16_t*>(mem);
        ring.holds.assign(ring.capacity, -1);
        ring.busy.assign(ring.capacity, 0);
    }
    FrameSource src;
    print_timestamp("START READ SAMPLES .csv");
    if (!src.open(o.input, o.inputFormat, W, H, o.bitDepth, ring.period)) return 1;
    if (ring.resident) {   // like the reference: every (distinct) frame is in memory before the device is touched
        for (int i = 0; i < ring.period; ++i)
            if (!src.next(ring.slot_ptr(i))) return 1;
    }
    print_timestamp("FINISH READ SAMPLES .csv");

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
    const size_t ncost = (size_t)sh.nCtus * MIP_COSTS_PER_CTU;
    // frame 0's tables for the reference's log, allocated and touched here like its return_* arrays (main.cpp:656-662)
    if (!o.noLog && !o.allFrames) {
        sh.keepCost.assign(ncost, 0);
        if (!o.compat) { sh.keepSad.assign(ncost, 0); sh.keepSatd.assign(ncost, 0); }
    }
    if (!o.digest.empty()) { sh.digCost.assign(o.nFrames, 0); sh.digMode.assign(o.nFrames, 0); sh.digBest.assign(o.nFrames, 0); }
    auto open_bin = [&](const std::string& path, const char* magic, uint32_t perCtu, const char* what) -> int {
        const int fd = open(path.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
        if (fd < 0) { perror(what); return -1; }
        uint32_t hdr[kBinHeader / 4] = {0};
        memcpy(hdr, magic, 8);
        hdr[2] = 1; hdr[3] = (uint32_t)W; hdr[4] = (uint32_t)H; hdr[5] = (uint32_t)o.nFrames; hdr[6] = (uint32_t)sh.nCtus;
        hdr[7] = perCtu; hdr[8] = (uint32_t)o.bitDepth; hdr[9] = (uint32_t)sh.filterType; hdr[10] = (uint32_t)o.kernelIdx; hdr[11] = (uint32_t)sh.k;
        if (pwrite(fd, hdr, sizeof(hdr), 0) != (ssize_t)sizeof(hdr)) { perror(what); close(fd); return -1; }
        return fd;
    };
    if (!o.binaryLog.empty() && (sh.binFd = open_bin(o.binaryLog, "MIPB200C", MIP_COSTS_PER_CTU, "error while opening the binary log")) < 0) { destroy_all(); return 1; }
    if (!o.compactLog.empty() && (sh.cmpFd = open_bin(o.compactLog, "MIPB200K", MIP_COMPACT_BYTES_PER_CTU, "error while opening the compact log")) < 0) { destroy_all(); return 1; }
    if (!o.decisionsBin.empty() && (sh.decFd = open_bin(o.decisionsBin, "MIPB200D", MIP_CUS_PER_CTU, "error while opening the binary decisions")) < 0) { destroy_all(); return 1; }
    if (!o.noLog && o.allFrames) {
        const std::string name = o.prefix + ".csv";
        if (!(sh.costLog.f = fopen(name.c_str(), "w"))) { perror("error while opening the output file"); destroy_all(); return 1; }
        fputs("POC,CTU,cuSizeName,W,H,CU,X,Y,Mode,SAD,SATD,minSadHad\n", sh.costLog.f);
    }
    if (!o.decisionsLog.empty()) {
        if (!(sh.decLog.f = fopen(o.decisionsLog.c_str(), "w"))) { perror("error while opening the decisions log"); destroy_all(); return 1; }
        std::string hdr = "POC,CTU,cuSizeName,W,H,CU,X,Y,BestMode,BestCost";
        for (int j = 2; j <= sh.k; ++j) hdr += ",Mode" + std::to_string(j) + ",Cost" + std::to_string(j);
        hdr += "\n";
        fputs(hdr.c_str(), sh.decLog.f);
    }
    // page-lock the ring so that every upload is a DMA from where the samples already are (no staging copy)
    ring.pinned = mipb200_pin_host_on(o.deviceIndex, ring.base, ring.fpx * sizeof(uint16_t) * (size_t)ring.capacity) == 0;
    std::thread reader;
    if (!ring.resident) {   // fill the ring before the window opens; from then on the reader refills released slots
        reader = std::thread(reader_thread, &sh, &src);
        std::unique_lock<std::mutex> lk(ring.mu);
        ring.cv.wait(lk, [&] { return ring.failed || ring.filled >= std::min<long>(ring.capacity, o.nFrames); });
    }
    printf("Frame ring: %d x %.1f MB %s, %s\n", ring.capacity, ring.fpx * 2 / 1e6, ring.pinned ? "page-locked" : "pageable",
           ring.resident ? "resident (every distinct frame loaded before the timed window)" : "streamed by a reader thread");
    print_timestamp("FINISH BUILD KERNELS");

    std::vector<unsigned long long> mj0(o.numGpus, 0), mj1(o.numGpus, 0);
    bool haveEnergy = o.energy;
    for (int g = 0; g < o.numGpus && haveEnergy; ++g)
        if (mipb200_device_energy_mj(o.deviceIndex + g, &mj0[g]) != 0) {
            printf("  [!] Energy counter unavailable: %s\n", mipb200_last_error());
            haveEnergy = false;
        }
    // ---- timed window: first upload -> last result consumed on the host (main.cpp:566-569, 1247-1250)
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
    if (reader.joinable()) { if (sh.errors) ring.fail(); reader.join(); }
    destroy_all();
    if (sh.binFd >= 0) close(sh.binFd);
    if (sh.decFd >= 0) close(sh.decFd);
    if (sh.cmpFd >= 0) close(sh.cmpFd);
    if (sh.costLog.f) fclose(sh.costLog.f);
    if (sh.decLog.f) fclose(sh.decLog.f);
    if (ring.pinned) mipb200_unpin_host(ring.base);
    free(ring.base);
    if (sh.errors) return 1;

    if (!o.noLog && !o.allFrames) {
        // like the reference, a log is written even when -l is omitted (file ".csv", main.cpp:1264-1269)
        std::string name = o.prefix + ".csv";
        FILE* f = fopen(name.c_str(), "w");
        if (!f) { perror("error while opening the output file"); return 1; }
        fputs("CTU,cuSizeName,W,H,CU,X,Y,Mode,SAD,SATD,minSadHad\n", f);
        // 13.2 M lines per 1080p frame: CTUs are formatted by all host threads into private buffers (one CTU = 4.4 MB of
        // text each) and written in CTU order
        const int32_t* sd = sh.keepSad.empty() ? nullptr : sh.keepSad.data();
        const int32_t* st = sh.keepSatd.empty() ? nullptr : sh.keepSatd.data();
        format_parallel(f, sh.nCtus, 6u << 20, [&](LogBuf& b, int ctu) { write_frame_log(b, 0, false, sh.keepCost.data(), sd, st, ctu, ctu + 1, W, o.compat); });
        fclose(f);
    }
    if (!o.digest.empty()) {
        FILE* f = fopen(o.digest.c_str(), "w");
        if (!f) { perror("error while opening the digest file"); return 1; }
        const bool haveCost = sh.emit & (MIPB200_EMIT_COSTS | MIPB200_EMIT_COSTS_COMPACT);
        fprintf(f, "POC,%sModes,BestCosts\n", (sh.emit & MIPB200_EMIT_COSTS_COMPACT) ? "CostsCompact," : haveCost ? "Costs," : "");
        for (int poc = 0; poc < o.nFrames; ++poc) {
            if (haveCost) fprintf(f, "%d,%016llx,%016llx,%016llx\n", poc, (unsigned long long)sh.digCost[poc], (unsigned long long)sh.digMode[poc], (unsigned long long)sh.digBest[poc]);
            else fprintf(f, "%d,%016llx,%016llx\n", poc, (unsigned long long)sh.digMode[poc], (unsigned long long)sh.digBest[poc]);
        }
        fclose(f);
    }

    // reportTimingResults_Compact (main_aux_functions.h:908-914)
    printf("=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=\n");
    printf("TIMING RESULTS (miliseconds)\n");
    printf("Elapsed time (ms) from writing samples to reading distortion (%dx), %d\n", o.nFrames, (int)lround(t1 - t0));
    printf("=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=\n\n");
    printf("Throughput: %.1f frames/s on %d GPU(s)\n", o.nFrames * 1e3 / (t1 - t0 > 0 ? t1 - t0 : 1e-3), o.numGpus);
    {
        struct rusage ru;
        if (getrusage(RUSAGE_SELF, &ru) == 0) printf("Peak host memory (MB), %.0f\n", ru.ru_maxrss / 1024.0);   // does not grow with -f: ring + engines' rings
    }
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
