// mip_work_list.h -- the fused kernel's work list, built once on the host.  Plain C++ (no CUDA): mip_kernels.cu copies the
// result to the device, tests/work_list_dump.cpp dumps it for tests/test_work_list.py, which checks it on a CPU against
// the CU tables (every (CU, mode) of a CTU exactly once, at the right place, in the right dispatch group).
//
// A CTU half's work is a list of warp tasks: 32 consecutive (CU, mode) pairs of one CU type (for the 64x64 type, whose
// pairs are few and long, four lanes per pair, a quarter of the strips each).  Per (half, task, lane) there is one 8-byte
// lane record; a frame's CTAs each take a contiguous chunk of a half's list (two splits: throughput / lone frame).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "mip_tables.h"

namespace mipb200 {

constexpr int TILE_ROWS = 64;           // a CTA works on the top or bottom half of a CTU
constexpr int MAX_CHUNKS = 64;
constexpr int MAX_WORK = 1700;          // rows of the lane-record table per CTU half; the last row is the end mark
constexpr int DEC_MAX = 2048;           // CUs per chunk the kernel's shared-memory argmin table can hold
constexpr int MAX_ORD = 2700;           // CUs per CTU half

// Lane record: .x = cuX (byte 0; bit 7 is never set) | cuY (byte 1) | mode (byte 2) | strip group << 24 | flags, .y = cost index
// in the CTU | decision slot << 17 | index of the shape inside its dispatch group << 29.  0xffffffff in .x = no more work.
// Fields sit on byte boundaries (one PRMT each), flags are tested in place, and the CU shape is a prefix code -- one flag
// bit per group, the most frequent group first -- instead of a number to be looked up in a 17-way switch.
struct LaneRec { uint32_t x, y; };
constexpr uint32_t REC_INRANGE = 1u << 26;     // the lane holds a real (CU, mode)
constexpr uint32_t REC_WRITER = 1u << 27;      // ... and is the one that writes its cost (strip group 0)
constexpr uint32_t REC_G64 = 1u << 28;         // 64x64 (12 modes, four lanes per (CU, mode))
constexpr uint32_t REC_GS1 = 1u << 29;         // 8x8, 16x4, 4x16, 32x4, 4x32 (16 modes), index in this order
constexpr uint32_t REC_GA32 = 1u << 30;        // 8x4, 4x8 (16 modes)
constexpr uint32_t REC_G4x4 = 1u << 31;        // 4x4 (32 modes)
                                               // no group flag: the eight other sizeId-2 shapes (12 modes), index = shape - S32x32

// the 17 distinct CU shapes
enum Shape {
    S64x64, S32x32, S32x16, S16x32, S32x8, S8x32, S16x16, S16x8, S8x16,  // sizeId 2
    S32x4, S4x32, S16x4, S4x16, S8x8, S8x4, S4x8,                        // sizeId 1
    S4x4,                                                                // sizeId 0
    NUM_SHAPES
};

inline int shape_of(int w, int h) {
    static const int tab[NUM_SHAPES][2] = {{64, 64}, {32, 32}, {32, 16}, {16, 32}, {32, 8}, {8, 32}, {16, 16}, {16, 8}, {8, 16},
                                           {32, 4},  {4, 32},  {16, 4},  {4, 16},  {8, 8},  {8, 4},  {4, 8},   {4, 4}};
    for (int i = 0; i < NUM_SHAPES; ++i)
        if (tab[i][0] == w && tab[i][1] == h) return i;
    return -1;
}

struct WorkList {
    std::vector<LaneRec> lanes[2];       // per half: 32 lane records per warp task
    std::vector<double> wcost[2];        // estimated cost of each warp task (for the split)
    std::vector<char> cut_ok[2];         // may a chunk boundary follow this warp task? (no CU's modes may be split)
    std::vector<int> ord_after[2];       // CU ordinal reached after this warp task (valid where cut_ok)
    uint16_t ord2cu[2][MAX_ORD];         // CU ordinal inside a half -> CU index inside the CTU (0..5379)
    int ord_total[2];                    // CUs per half
    uint8_t shape[MIP_NUM_TYPES], parts_log2[MIP_NUM_TYPES];
    // the two splits of each half's list into chunks: [split][half][chunk]
    int chunks[2];
    int begin[2][2][MAX_CHUNKS + 1];           // first warp task of each chunk
    uint16_t chunk_ord[2][2][MAX_CHUNKS + 1];  // first CU ordinal of each chunk
};

// The lane records of both halves.  false = a CU table this kernel's assumptions do not hold for.
inline bool build_work_list(WorkList& wl) {
    wl.ord_total[0] = wl.ord_total[1] = 0;
    for (int hf = 0; hf < 2; ++hf) { wl.lanes[hf].clear(); wl.wcost[hf].clear(); wl.cut_ok[hf].clear(); wl.ord_after[hf].clear(); }
    for (int t = 0; t < MIP_NUM_TYPES; ++t) {
        const mip_cu_type_t& s = MIP_TYPES[t];
        const int shape = shape_of(s.w, s.h);
        if (shape < 0) return false;
        wl.shape[t] = (uint8_t)shape;
        const int pl2 = wl.parts_log2[t] = (s.w == 64 && s.h == 64) ? 2 : 0;
        // CUs per CTU half: CU order is raster, so each half is one contiguous run; no CU crosses y = 64
        for (int hf = 0; hf < 2; ++hf) {
            int first = -1, cnt = 0;
            for (int cu = 0; cu < s.n; ++cu) {
                const int y = s.ys[cu / s.cols];
                if (y / 64 != (y + s.h - 1) / 64) return false;
                if (y / 64 == hf) { if (first < 0) first = cu; else if (cu != first + cnt) return false; cnt++; }
            }
            const int first_cu = first < 0 ? 0 : first;
            const double mv = s.size_id == 2 ? 700.0 : (s.size_id == 1 ? 200.0 : 120.0);
            const double c = (mv + 11.0 * s.w * s.h + 2.0 * (s.w + s.h) + 60.0) / (1 << pl2) + (pl2 ? mv : 0.0);
            const int per_cu = s.modes << pl2, ntask = cnt * per_cu;
            const int nw = (ntask + 31) / 32;
            for (int w = 0; w < nw; ++w) {
                for (int lane = 0; lane < 32; ++lane) {
                    const int task = w * 32 + lane, in_range = task < ntask, tcl = in_range ? task : ntask - 1;
                    const int part = tcl & ((1 << pl2) - 1), cm = tcl >> pl2;
                    const int cu_local = cm / s.modes, mode = cm % s.modes, cu = first_cu + cu_local;
                    const int cx = s.xs[cu % s.cols], cy = s.ys[cu / s.cols] - hf * TILE_ROWS;
                    const uint32_t coff = s.cost_off + (uint32_t)cu * s.modes + mode, slot = (uint32_t)(wl.ord_total[hf] + cu_local);
                    if (cx > 127 || cy < 0 || cy > 63 || mode > 31 || part > 3 || coff >= (1u << 17) || slot >= (1u << 12)) return false;
                    // dispatch group (one flag bit each, most frequent first in the kernel) and the shape's index inside it
                    uint32_t grp = 0, sub = 0;
                    switch (shape) {
                        case S4x4: grp = REC_G4x4; break;
                        case S8x4: grp = REC_GA32; sub = 0; break;
                        case S4x8: grp = REC_GA32; sub = 1; break;
                        case S8x8: grp = REC_GS1; sub = 0; break;
                        case S16x4: grp = REC_GS1; sub = 1; break;
                        case S4x16: grp = REC_GS1; sub = 2; break;
                        case S32x4: grp = REC_GS1; sub = 3; break;
                        case S4x32: grp = REC_GS1; sub = 4; break;
                        case S64x64: grp = REC_G64; break;
                        default: sub = (uint32_t)shape - S32x32; break;     // the eight other sizeId-2 shapes
                    }
                    // the kernel's decision code relies on: 32 modes = one CU per task (no lane past the type's end), 16 modes = one CU per half task
                    if (sub > 7 || s.modes != (grp == REC_G4x4 ? 32 : (grp & (REC_GA32 | REC_GS1)) ? 16 : 12) || (grp == REC_G4x4 && !in_range)) return false;
                    wl.lanes[hf].push_back(LaneRec{(uint32_t)cx | ((uint32_t)cy << 8) | ((uint32_t)mode << 16) | ((uint32_t)part << 24) | (in_range ? REC_INRANGE : 0u) |
                                                       (in_range && part == 0 ? REC_WRITER : 0u) | grp,
                                                   coff | (slot << 17) | (sub << 29)});
                }
                wl.wcost[hf].push_back(c);
                const int done = std::min(ntask, 32 * (w + 1));
                wl.cut_ok[hf].push_back(done % per_cu == 0);
                wl.ord_after[hf].push_back(wl.ord_total[hf] + done / per_cu);
            }
            if (wl.ord_total[hf] + cnt > MAX_ORD) return false;
            for (int k = 0; k < cnt; ++k) wl.ord2cu[hf][wl.ord_total[hf] + k] = (uint16_t)(s.cu_off + first_cu + k);
            wl.ord_total[hf] += cnt;
        }
    }
    for (int hf = 0; hf < 2; ++hf)
        if ((int)wl.wcost[hf].size() > MAX_WORK - 2) return false;      // the table's last row is the end mark
    return true;
}

// Contiguous, cost-balanced partition of each half's list into chunks, once per split: nchunks[sp] chunks with relative
// cost shares weights[sp][0..] (equal shares where weights[sp] is null).  false = a chunk holds more CUs than the kernel's
// shared-memory decision table (use more chunks).
inline bool split_work_list(WorkList& wl, const int* nchunks, const double* const* weights) {
    for (int sp = 0; sp < 2; ++sp) {
        const int chunks = wl.chunks[sp] = std::min(std::max(nchunks[sp], 1), MAX_CHUNKS);
        // cum[k] = share of a half's cost that lies before chunk k
        double cum[MAX_CHUNKS + 1], wsum = 0;
        for (int k = 0; k < chunks; ++k) wsum += weights[sp] ? weights[sp][k] : 1.0;
        cum[0] = 0;
        for (int k = 0; k < chunks; ++k) cum[k + 1] = cum[k] + (weights[sp] ? weights[sp][k] : 1.0) / wsum;
        for (int hf = 0; hf < 2; ++hf) {
            const size_t ntasks = wl.wcost[hf].size();
            double total = 0;
            for (double c : wl.wcost[hf]) total += c;
            wl.begin[sp][hf][0] = 0;
            wl.chunk_ord[sp][hf][0] = 0;
            double acc = 0;
            int k = 1;
            for (size_t i = 0; i < ntasks && k < chunks; ++i) {
                acc += wl.wcost[hf][i];
                if (acc >= total * cum[k] && wl.cut_ok[hf][i]) { wl.chunk_ord[sp][hf][k] = (uint16_t)wl.ord_after[hf][i]; wl.begin[sp][hf][k++] = (int)i + 1; }
            }
            while (k <= MAX_CHUNKS) { wl.chunk_ord[sp][hf][k] = (uint16_t)wl.ord_total[hf]; wl.begin[sp][hf][k++] = (int)ntasks; }
            for (int q = 0; q < chunks; ++q)
                if (wl.chunk_ord[sp][hf][q + 1] - wl.chunk_ord[sp][hf][q] > DEC_MAX) return false;
        }
    }
    return true;
}

}  // namespace mipb200
