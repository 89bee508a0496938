/*
 * mip_oracle.c -- CPU restatement of the reference's MIP mode-decision path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, link or call it.  The
 * product path (vvc-mip-gpu_b200/csrc) never does and fails loudly without its CUDA library.
 *
 * It restates, stage by stage, what the reference's OpenCL kernels compute for one frame
 * (citations are file:line into the reference repository iagostorch/VVC-MIP-GPU):
 *
 *   filterFrame_*        intra.cl:1639-3823   -> mipo_filter_frame()
 *   initBoundaries       intra.cl:17-344      -> cu_boundaries()
 *   MIP_ReducedPred      intra.cl:349-543     -> reduced_prediction()
 *   upsampleDistortion   intra.cl:545-1171    -> upsample(), cu_distortion()
 *   satd_4x4             kernel_aux_functions.cl:142-249 -> satd4x4()
 *
 * Output layout = the reference's minSadHad buffer: cost[ctu][97840] with the per-type
 * offsets of ALL_stridedDistortionsPerCtu (constants.h:1558-1631) and cu*modes + mode inside
 * a type (intra.cl:1144-1148).  Values are int32 (the reference's `long` holds < 2^24).
 *
 * Parity definition (reference defects make anything else meaningless):
 *   - frame-0 semantics for every frame;
 *   - CUs that do not lie fully inside the frame (Y + h > H) are skipped by the reference
 *     (intra.cl:96-98, 232-234, 717) and left as garbage; here they get MIPO_SKIPPED (-1);
 *   - the four 1-D 3x3/5x5 filters use the intended fill-then-fetch order, and samples
 *     beyond the frame bottom count as 0 for the 1-D 3x3 type (intra.cl:3330-3332 reads
 *     them unguarded from a fresh buffer).
 *
 * Pin status: the reference ships no golden vectors and cannot run in the build container
 * (OpenCL only).  The oracle is pinned against (1) the survey's known-answer vectors
 * (tests/golden/survey_kat.json), (2) outputs of the reference's own, unmodified OpenCL
 * kernels run on a B200 through NVIDIA's OpenCL driver by oracle/ocl_ref (fixtures under
 * tests/golden/, generating scripts committed): six complete cost tables of small frames, all
 * 32 filter outputs, and per-CTU hashes of the complete tables of three 1080p frames and one
 * 2160p frame, and (3) the reference's own CPU filters compiled from its checkout
 * (oracle/cpu_ref).  See tests/test_oracle_golden.py and tests/test_oracle_cpu_ref.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../vvc-mip-gpu_b200/csrc/mip_filters.h"
#include "../vvc-mip-gpu_b200/csrc/mip_matrices.h"
#include "../vvc-mip-gpu_b200/csrc/mip_tables.h"

#define MIPO_SKIPPED (-1)
#define API __attribute__((visibility("default")))

/* Sample bit depth.  The reference hard-wires 10 bits (valueDC 1 << 9 intra.cl:61, 1 << 9 intra.cl:446, clamp to 1023
 * intra.cl:482); 8 and 12 generalise those three constants the way the VVC specification does (1 << (bitDepth - 1),
 * (1 << bitDepth) - 1).  Parity with the reference exists only for 10.  Set once before a run; not thread-safe. */
static int g_bit_depth = 10;
API int mipo_set_bit_depth(int bits) {
    if (bits != 8 && bits != 10 && bits != 12) return -1;
    g_bit_depth = bits;
    return 0;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

/* ------------------------------------------------------------------------------------------
 * Filters ("alternative samples").  All eight CLI filter kernels reduce to four integer
 * functions {1-D,2-D} x {3x3,5x5}: the float kernels accumulate integer-valued fp32 numbers
 * and finish with round(N/S), which equals (N + S/2)/S for every N, S that can occur
 * (checked exhaustively in tests/test_oracle_units.py::test_float_rounding_equals_int).
 *
 * filter_type: 1 1d_int, 2 1d_float, 3 2d_int_quarterCtu, 4 2d_float_quarterCtu,
 *              5 1d_int_5x5, 6 1d_float_5x5, 7 2d_int_5x5_quarterCtu, 8 2d_float_5x5_quarterCtu
 *              (order of availableFilters, constants.h:25-34)
 * ---------------------------------------------------------------------------------------- */
static inline int sample_or_zero(const uint16_t* f, int W, int H, int x, int y) {
    return (x >= 0 && x < W && y >= 0 && y < H) ? f[(size_t)y * W + x] : 0;
}

/* 2-D, R = 1 (3x3) or 2 (5x5): numerator and denominator over the in-frame taps only
 * (intra.cl:2993-3011 int 3x3, 1776-1794 float 3x3, 3215-3235 int 5x5, 2487-2507 float 5x5) */
static int filt_2d(const uint16_t* f, int W, int H, int x, int y, int R, int kidx) {
    int num = 0, den = 0;
    for (int dy = -R; dy <= R; ++dy)
        for (int dx = -R; dx <= R; ++dx) {
            int xx = x + dx, yy = y + dy;
            if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
            int k = (R == 1) ? mip_k3(kidx, dy, dx) : mip_k5(kidx, dy, dx);
            num += k * f[(size_t)yy * W + xx];
            den += k;
        }
    return (num + den / 2) / den;
}

/* 1-D 3x3: horizontal then vertical 3-tap pass with the FIRST ROW of the 3x3 table
 * (intra.cl:3274-3278); missing taps contribute 0; the denominator comes from the position
 * class, not from the taps (intra.cl:3281-3285, 3437-3466; float: 1848-1852, 2017-2046). */
static int filt_1d_3(const uint16_t* f, int W, int H, int x, int y, int kidx) {
    int k[3] = {mip_k3(kidx, -1, -1), mip_k3(kidx, -1, 0), mip_k3(kidx, -1, 1)};
    int num = 0;
    for (int dy = -1; dy <= 1; ++dy) {
        int hor = 0;
        for (int dx = -1; dx <= 1; ++dx) hor += k[dx + 1] * sample_or_zero(f, W, H, x + dx, y + dy);
        num += k[dy + 1] * hor;
    }
    int full = 4 * k[0] + 4 * k[1] + k[1] * k[1];
    int corner = 1 * k[0] + 2 * k[1] + k[1] * k[1];
    int edge = 2 * k[0] + 3 * k[1] + k[1] * k[1];
    int nEdges = (x == 0) + (x == W - 1) + (y == 0) + (y == H - 1);
    int den = nEdges >= 2 ? corner : (nEdges == 1 ? edge : full);
    return (num + den / 2) / den;
}

/* 1-D 5x5: separable 5-tap passes with the first row of the 5x5 table (intra.cl:3515-3521),
 * denominator by position class computed from the 2-D table (intra.cl:3523-3551, 3769-3788;
 * float: 2554-2582, 2800-2819). */
static int ksum5(int kidx, int i0, int j0) {
    int s = 0;
    for (int i = i0; i < 5; ++i)
        for (int j = j0; j < 5; ++j) s += mip_k5(kidx, i - 2, j - 2);
    return s;
}
static int filt_1d_5(const uint16_t* f, int W, int H, int x, int y, int kidx) {
    int k[5];
    for (int i = 0; i < 5; ++i) k[i] = mip_k5(kidx, -2, i - 2);
    int num = 0;
    for (int dy = -2; dy <= 2; ++dy) {
        int hor = 0;
        for (int dx = -2; dx <= 2; ++dx) hor += k[dx + 2] * sample_or_zero(f, W, H, x + dx, y + dy);
        num += k[dy + 2] * hor;
    }
    int outerTB = (y == 0) || (y == H - 1), innerTB = (y == 1) || (y == H - 2);
    int outerLR = (x == 0) || (x == W - 1), innerLR = (x == 1) || (x == W - 2);
    int outerCorner = outerTB && outerLR, innerCorner = innerTB && innerLR;
    int iface = (outerLR && innerTB) || (innerLR && outerTB);
    int outerEdge = !outerCorner && !iface && (outerTB || outerLR);
    int innerEdge = !innerCorner && !iface && (innerTB || innerLR);
    /* start from the tap-corrected scale of the vertical pass (intra.cl:3753-3758) ... */
    int den = ksum5(kidx, 0, 0);
    for (int dy = -2; dy <= 2; ++dy)
        if (y + dy < 0 || y + dy >= H) den -= k[dy + 2];
    /* ... which every position class then overrides (intra.cl:3782-3786) */
    if (outerCorner) den = ksum5(kidx, 2, 2);
    if (innerCorner) den = ksum5(kidx, 1, 1);
    if (outerEdge) den = ksum5(kidx, 0, 2);
    if (innerEdge) den = ksum5(kidx, 0, 1);
    if (iface) den = ksum5(kidx, 1, 2);
    return (num + den / 2) / den;
}

API int mipo_filter_frame(const uint16_t* frame, int W, int H, int filter_type, int kernel_idx, uint16_t* out) {
    if (filter_type < 1 || filter_type > 8) return -1;
    int is5 = filter_type >= 5;
    int is2d = (filter_type == 3 || filter_type == 4 || filter_type == 7 || filter_type == 8);
    if (kernel_idx < 0 || kernel_idx >= (is5 ? MIP_NUM_K5 : MIP_NUM_K3)) return -2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int v;
            if (is2d) v = filt_2d(frame, W, H, x, y, is5 ? 2 : 1, kernel_idx);
            else if (is5) v = filt_1d_5(frame, W, H, x, y, kernel_idx);
            else v = filt_1d_3(frame, W, H, x, y, kernel_idx);
            out[(size_t)y * W + x] = (uint16_t)v;
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * initBoundaries: complete (refT/refL) and reduced (redT/redL) boundaries of one CU.
 * F = frame the references come from (original or filtered), X,Y absolute CU origin.
 * ---------------------------------------------------------------------------------------- */
static void cu_boundaries(const uint16_t* F, int W, int X, int Y, int w, int h, int b,
                          int* refT, int* refL, int* redT, int* redL) {
    const int valueDC = 1 << (g_bit_depth - 1); /* intra.cl:61 */
    for (int i = 0; i < w; ++i) {  /* intra.cl:96-107 */
        if (Y > 0) refT[i] = F[(size_t)(Y - 1) * W + X + i];
        else if (X == 0) refT[i] = valueDC;
        else refT[i] = F[X - 1];
    }
    for (int j = 0; j < h; ++j) {  /* intra.cl:232-243 */
        if (X > 0) refL[j] = F[(size_t)(Y + j) * W + X - 1];
        else if (Y == 0) refL[j] = valueDC;
        else refL[j] = F[(size_t)(Y - 1) * W];
    }
    /* intra.cl:71-73, 127-141 (top) and 202-204, 260-279 (left).  With a down-sampling
     * factor of 1 the reference's rounding offset 1 << -1 truncates to 0 in a short. */
    int dT = w / b, lT = ilog2(dT), rT = dT > 1 ? (1 << (lT - 1)) : 0;
    for (int q = 0; q < b; ++q) {
        int s = 0;
        for (int t = 0; t < dT; ++t) s += refT[q * dT + t];
        redT[q] = (s + rT) >> lT;
    }
    int dL = h / b, lL = ilog2(dL), rL = dL > 1 ? (1 << (lL - 1)) : 0;
    for (int q = 0; q < b; ++q) {
        int s = 0;
        for (int t = 0; t < dL; ++t) s += refL[q * dL + t];
        redL[q] = (s + rL) >> lL;
    }
}

/* ------------------------------------------------------------------------------------------
 * MIP_ReducedPred for one (CU, mode): r x r reduced prediction (intra.cl:415-487).
 * ---------------------------------------------------------------------------------------- */
static void reduced_prediction(int size_id, int mode, const int* redT, const int* redL, int* red) {
    const int M = size_id == 2 ? 6 : (size_id == 1 ? 8 : 16);
    const int r = size_id == 2 ? 8 : 4;
    const int b = size_id == 0 ? 2 : 4;
    const int tr = mode >= M, mat = mode % M;  /* intra.cl:417-418 */
    int in[8];
    for (int i = 0; i < b; ++i) {  /* intra.cl:434, 440 */
        in[i] = tr ? redL[i] : redT[i];
        in[b + i] = tr ? redT[i] : redL[i];
    }
    const int first = in[0];
    for (int i = 0; i < 2 * b; ++i) in[i] -= first;          /* intra.cl:445 */
    in[0] = (size_id == 2) ? 0 : (1 << (g_bit_depth - 1)) - first; /* intra.cl:446 */
    int sum = 0;
    for (int i = 0; i < 2 * b; ++i) sum += in[i];             /* intra.cl:449-452 */
    const int offset = (1 << 5) - 32 * sum;                   /* intra.cl:454 */
    for (int p = 0; p < r * r; ++p) {
        int v = offset;
        for (int i = 0; i < 2 * b; ++i) {                    /* intra.cl:474-479 */
            const int c = size_id == 2 ? mip_mat_id2(mat, p, i) : (size_id == 1 ? mip_mat_id1(mat, p, i) : mip_mat_id0(mat, p, i));
            v += c * in[i];
        }
        v = (v >> 6) + first;                                 /* intra.cl:481 */
        v = iclamp(v, 0, (1 << g_bit_depth) - 1);               /* intra.cl:482 */
        int x = p % r, y = p / r;
        red[tr ? (x * r + y) : p] = v;                        /* intra.cl:485-487 */
    }
}

/* ------------------------------------------------------------------------------------------
 * upsampleDistortion, part 1: bilinear up-sampling of the reduced prediction to w x h
 * (intra.cl:645-652, 816-893).  sizeId 0 copies (intra.cl:726-727).
 * ---------------------------------------------------------------------------------------- */
static void upsample(const int* red, int r, int w, int h, const int* refT, const int* refL, int* pred) {
    const int uH = w / r, uV = h / r;
    const int lH = ilog2(uH), lV = ilog2(uV);
    const int rH = uH > 1 ? (1 << (lH - 1)) : 0, rV = uV > 1 ? (1 << (lV - 1)) : 0;
    /* horizontal pass on the rows that hold reduced samples (intra.cl:818-844) */
    for (int j = 0; j < r; ++j) {
        int y = j * uV + uV - 1;
        for (int x = 0; x < w; ++x) {
            int o = x % uH + 1;
            int after = red[j * r + (x >> lH)];
            int before = (x < uH) ? refL[y] : red[j * r + (x >> lH) - 1];
            pred[y * w + x] = ((uH - o) * before + o * after + rH) >> lH;
        }
    }
    /* vertical pass over all rows (intra.cl:869-893); rows y % uV == uV-1 reproduce themselves */
    for (int y = 0; y < h; ++y) {
        if (y % uV == uV - 1) continue;
        int y0 = (y >> lV) << lV, o = y % uV + 1;
        for (int x = 0; x < w; ++x) {
            int after = pred[(y0 + uV - 1) * w + x];
            int before = (y < uV) ? refT[x] : pred[(y0 - 1) * w + x];
            pred[y * w + x] = ((uV - o) * before + o * after + rV) >> lV;
        }
    }
}

/* satd_4x4 (kernel_aux_functions.cl:142-249): VTM xCalcHADs4x4 with the mean-scaled DC term */
static int satd4x4(const int* diff /* 16 values, raster order */) {
    int m[16], d[16];
    for (int i = 0; i < 4; ++i) {  /* 1st + 2nd stage: vertical butterflies (:167-200) */
        int a0 = diff[i] + diff[12 + i], a1 = diff[4 + i] + diff[8 + i];
        int a2 = diff[4 + i] - diff[8 + i], a3 = diff[i] - diff[12 + i];
        m[i] = a0 + a1; m[4 + i] = a2 + a3; m[8 + i] = a0 - a1; m[12 + i] = a3 - a2;
    }
    for (int r = 0; r < 4; ++r) {  /* 3rd + 4th stage: horizontal butterflies (:203-236) */
        const int* q = m + 4 * r;
        int b0 = q[0] + q[3], b1 = q[1] + q[2], b2 = q[1] - q[2], b3 = q[0] - q[3];
        d[4 * r + 0] = b0 + b1; d[4 * r + 1] = b0 - b1; d[4 * r + 2] = b2 + b3; d[4 * r + 3] = b3 - b2;
    }
    int satd = 0;
    for (int k = 0; k < 16; ++k) satd += abs(d[k]);  /* :238-241 */
    satd -= abs(d[0]);                               /* :244 */
    satd += abs(d[0]) >> 2;                          /* :245 */
    return (satd + 1) >> 1;                          /* :246 */
}

/* upsampleDistortion, part 2: SAD, SATD over all 4x4 sub-blocks, min(2*SAD, SATD)
 * (intra.cl:922-1053, 1166) */
static void cu_distortion(const uint16_t* O, int W, int X, int Y, int w, int h, const int* pred,
                          int* sad_out, int* satd_out) {
    int sad = 0, satd = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) sad += abs((int)O[(size_t)(Y + y) * W + X + x] - pred[y * w + x]);
    for (int by = 0; by < h; by += 4)
        for (int bx = 0; bx < w; bx += 4) {
            int diff[16];
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j)
                    diff[4 * i + j] = (int)O[(size_t)(Y + by + i) * W + X + bx + j] - pred[(by + i) * w + bx + j];
            satd += satd4x4(diff);
        }
    *sad_out = sad;
    *satd_out = satd;
}

/* One (CTU, CU type): every CU, every mode. */
static void ctu_type_costs(const uint16_t* O, const uint16_t* F, int W, int H, int ctuX, int ctuY, int t,
                           int32_t* cost, int32_t* sad_o, int32_t* satd_o) {
    const mip_cu_type_t* ty = &MIP_TYPES[t];
    const int w = ty->w, h = ty->h, sid = ty->size_id, modes = ty->modes;
    const int r = sid == 2 ? 8 : 4, b = sid == 0 ? 2 : 4;
    int refT[64], refL[64], redT[4], redL[4], red[64];
    int* pred = (int*)malloc(sizeof(int) * 64 * 64);
    for (int cu = 0; cu < ty->n; ++cu) {
        int X = ctuX + ty->xs[cu % ty->cols], Y = ctuY + ty->ys[cu / ty->cols];
        int32_t* c = cost + ty->cost_off + cu * modes;
        int32_t* s1 = sad_o ? sad_o + ty->cost_off + cu * modes : NULL;
        int32_t* s2 = satd_o ? satd_o + ty->cost_off + cu * modes : NULL;
        if (Y + h > H || X + w > W) {  /* bottom: intra.cl:96, 232, 717 -- skipped by the reference, garbage there;
                                        * right: the reference has no x-guards at all (it wraps into the next row for the
                                        * whitelisted 832x480 / 416x240), so "skipped" is this project's definition */
            for (int m = 0; m < modes; ++m) {
                c[m] = MIPO_SKIPPED;
                if (s1) s1[m] = MIPO_SKIPPED;
                if (s2) s2[m] = MIPO_SKIPPED;
            }
            continue;
        }
        cu_boundaries(F, W, X, Y, w, h, b, refT, refL, redT, redL);
        for (int m = 0; m < modes; ++m) {
            int sad, satd;
            reduced_prediction(sid, m, redT, redL, red);
            if (sid == 0) memcpy(pred, red, sizeof(int) * 16);
            else upsample(red, r, w, h, refT, refL, pred);
            cu_distortion(O, W, X, Y, w, h, pred, &sad, &satd);
            c[m] = imin(2 * sad, satd);  /* intra.cl:1166 */
            if (s1) s1[m] = sad;
            if (s2) s2[m] = satd;
        }
    }
    free(pred);
}

/* All costs of one frame.  orig = samples the distortion is measured against
 * (main.cpp:1017), ref = samples the boundaries are taken from (orig, or the filtered frame
 * when alternative samples are on, main.cpp:818-822).  cost/sad/satd: [nCTU][97840] int32;
 * sad/satd may be NULL.  threads <= 0 -> all cores. */
API int mipo_frame_costs(const uint16_t* orig, const uint16_t* ref, int W, int H,
                         int32_t* cost, int32_t* sad, int32_t* satd, int threads) {
    if (W <= 0 || H <= 0 || W % 4 != 0 || H % 4 != 0) return -1;
    const int cols = (W + 127) / 128, rows = (H + 127) / 128, nctu = cols * rows;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int job = 0; job < nctu * MIP_NUM_TYPES; ++job) {
        int ctu = job / MIP_NUM_TYPES, t = job % MIP_NUM_TYPES;
        size_t base = (size_t)ctu * MIP_COSTS_PER_CTU;
        ctu_type_costs(orig, ref, W, H, 128 * (ctu % cols), 128 * (ctu / cols), t, cost + base,
                       sad ? sad + base : NULL, satd ? satd + base : NULL);
    }
    return 0;
}

/* filter (optional) + costs: what one iteration of the reference's frame loop produces
 * (main.cpp:678-1241). */
API int mipo_run_frame(const uint16_t* frame, int W, int H, int filter_type, int kernel_idx,
                       int32_t* cost, int32_t* sad, int32_t* satd, int threads) {
    if (filter_type == 0) return mipo_frame_costs(frame, frame, W, H, cost, sad, satd, threads);
    uint16_t* filt = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)W * H);
    if (!filt) return -3;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int rc = mipo_filter_frame(frame, W, H, filter_type, kernel_idx, filt);
    if (rc == 0) rc = mipo_frame_costs(frame, filt, W, H, cost, sad, satd, threads);
    free(filt);
    return rc;
}

/* Decision derived from the costs: per CU argmin over modes, lowest mode wins ties.
 * best_mode[nCTU][5380] (0xFF for skipped CUs), best_cost[nCTU][5380]. */
API void mipo_decisions(const int32_t* cost, int nctu, uint8_t* best_mode, int32_t* best_cost) {
    for (int ctu = 0; ctu < nctu; ++ctu)
        for (int t = 0; t < MIP_NUM_TYPES; ++t) {
            const mip_cu_type_t* ty = &MIP_TYPES[t];
            for (int cu = 0; cu < ty->n; ++cu) {
                const int32_t* c = cost + (size_t)ctu * MIP_COSTS_PER_CTU + ty->cost_off + cu * ty->modes;
                int bm = 0xFF, bc = MIPO_SKIPPED;
                if (c[0] != MIPO_SKIPPED) {
                    bm = 0; bc = c[0];
                    for (int m = 1; m < ty->modes; ++m)
                        if (c[m] < bc) { bc = c[m]; bm = m; }
                }
                best_mode[(size_t)ctu * MIP_CUS_PER_CTU + ty->cu_off + cu] = (uint8_t)bm;
                best_cost[(size_t)ctu * MIP_CUS_PER_CTU + ty->cu_off + cu] = bc;
            }
        }
}

/* unit-test hooks */
API int mipo_satd4x4(const int* diff16) { return satd4x4(diff16); }
API void mipo_reduced_prediction(int size_id, int mode, const int* redT, const int* redL, int* red) {
    reduced_prediction(size_id, mode, redT, redL, red);
}
API void mipo_upsample(const int* red, int r, int w, int h, const int* refT, const int* refL, int* pred) {
    upsample(red, r, w, h, refT, refL, pred);
}
API void mipo_cu_boundaries(const uint16_t* F, int W, int X, int Y, int w, int h, int b,
                            int* refT, int* refL, int* redT, int* redL) {
    cu_boundaries(F, W, X, Y, w, h, b, refT, refL, redT, redL);
}
