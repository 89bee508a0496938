// mip_kernels.cu -- sm_100a kernels of the MIP mode-decision engine.
//
// What the reference does in four global-memory-coupled OpenCL kernels
// (initBoundaries -> MIP_ReducedPred -> upsampleDistortion x3, intra.cl:17-1171, ~1.4 GB of
// intermediate traffic per 1080p frame) is done here in ONE kernel that never leaves the SM:
//
//   CTA  = one (CTU half, chunk of that half's work).  No CU crosses the y = 64 line of its CTU, so
//          the top and bottom 128x64 halves are independent.  The CTA stages the half's original
//          samples (as int32 holding o+1, so a 4-pixel row of a 4x4 block is one LDS.128) and the
//          reference-sample tile with its top/left halo in shared memory; two CTAs share an SM.
//          The tile arrives as ONE TMA box (cp.async.bulk.tensor.2d, 144x69 uint16 incl. a 3-sample halo,
//          out-of-frame samples zero-filled by the hardware) signalled on an mbarrier; the CTA then
//          expands it to the int32 original tile and to the reference tile -- which is where the
//          "alternative samples" low-pass filter (intra.cl:1639-3823) is applied, so filtered references
//          never exist in HBM.
//   lane = one (CU, mode) pair = one output cost.  32 consecutive (CU, mode) pairs of one CU
//          type form a warp task, so all lanes run the same shape-specialised code, nothing
//          is reduced across lanes, there is no barrier after staging, and the 32 costs of a
//          warp leave as one coalesced 128-byte store in the reference's buffer order.  Tasks are
//          drawn from a shared-memory counter one ahead; what each lane does in a task is a
//          ready-made 8-byte record of an L2-resident table (g_lane) built once on the host.
//   per lane: reduced boundaries (A.2) -> matrix-vector product with IDP.2A on (coef-32)
//          signed bytes (A.3) into a private shared-memory column -> strip-wise bilinear
//          up-sampling (A.4) -> per 4x4 block: difference, SAD, Hadamard SATD (A.5).
//
// (A.x = arithmetic specification in SURVEY.md Appendix A; file:line = reference repository.)
#include "mip_kernels.h"

#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "mip_compact.h"
#include "mip_filters.h"
#include "mip_matrices.h"
#include "mip_tables.h"
#include "mip_work_list.h"

namespace mipb200 {

// ------------------------------------------------------------------------------------------
// Geometry / layout constants
// ------------------------------------------------------------------------------------------
#ifndef MIP_NT
#define MIP_NT 384
#endif
constexpr int NT = MIP_NT;              // threads per CTA
constexpr int OS = 132;                 // s_orig row stride in int32 words (128 + 4: rows shift 4 banks)
// Reference samples are only ever read on the row above / the column left of a CU, and every CU origin is a multiple
// of 4: the reference tile keeps 18 row slots (halo row -1, rows 3,7,..,63, frame row 0) and 34 column slots (halo
// column -1, columns 3,7,..,127, frame column 0), the columns stored transposed so that a left boundary is contiguous.
constexpr int RT_STRIDE = 136, RT_SLOTS = 18, RT_ROW0 = 17;   // row slot: [x + 8], x = -8..127
constexpr int RL_STRIDE = 72, RL_SLOTS = 34, RL_COL0 = 33;    // column slot: [y + 8], y = -8..63
constexpr int RED_WORDS = 32;           // packed reduced-prediction words per thread (64 samples)

// Every matrix is kept twice: as is, and with its rows permuted by the transposition p = y*r + x -> x*r + y, so that
// transposed modes read row "output position" like the others.  The lanes of a warp hold different modes, i.e. read
// different matrices at the same row: the per-matrix pads spread the matrices over distinct banks, and each permuted
// copy starts a whole number of bank groups after the plain one (48 B / 64 B / 64 B modulo 128), so plain and transposed
// lanes never collide either.
constexpr int M2_STRIDE = 528, M1_STRIDE = 144, M0_STRIDE = 80;   // 64x8 B + 16, 16x8 B + 16, 16x4 B + 16
constexpr int M2_OFF = 0, M2T_OFF = M2_OFF + 6 * M2_STRIDE;       // 3168 = 24*128 + 96
constexpr int M1_OFF = M2T_OFF + 6 * M2_STRIDE, M1T_OFF = M1_OFF + 8 * M1_STRIDE;    // +1152 = 9*128
constexpr int M0_OFF = M1T_OFF + 8 * M1_STRIDE, M0T_OFF = M0_OFF + 16 * M0_STRIDE;   // +1280
constexpr int MAT_BYTES = M0T_OFF + 16 * M0_STRIDE;   // 11200
// sizeId 2 / 1: two 8-byte rows per LDS.128, so matrices start 16-byte aligned and 4 banks apart (6 + 6 or 8 + 8 matrices
// of a warp = 12 or 16 distinct 16-byte rows = the minimum of 2 wavefronts); sizeId 0: four 4-byte rows per LDS.128, the
// 16 + 16 matrices of a warp 20 banks apart = 8 distinct 4-bank groups used by 4 matrices each = the minimum of 4 wavefronts.
static_assert(M2_STRIDE % 16 == 0 && M1_STRIDE % 16 == 0 && M0_STRIDE % 16 == 0 && M2T_OFF % 16 == 0 && M1_OFF % 16 == 0 &&
              M1T_OFF % 16 == 0 && M0_OFF % 16 == 0 && M0T_OFF % 16 == 0, "LDS.128 alignment");
static_assert(M2T_OFF == M2_OFF + 6 * M2_STRIDE && M1T_OFF == M1_OFF + 8 * M1_STRIDE && M0T_OFF == M0_OFF + 16 * M0_STRIDE, "transposed copies follow the plain matrices");
static_assert((M2_STRIDE / 4) % 32 == 4 && (M1_STRIDE / 4) % 32 == 4 && (M0_STRIDE / 4) % 8 == 4, "bank phases");

constexpr int SM_RED = 0;                                      // first: the TMA box lands here (128-byte aligned)
constexpr int SM_RED_BYTES = RED_WORDS * NT * 4;               // 49152 at NT = 384
constexpr int SM_ORIG = SM_RED + SM_RED_BYTES;
constexpr int SM_ORIG_BYTES = TILE_ROWS * OS * 4;              // 33792
constexpr int SM_REF = SM_ORIG + SM_ORIG_BYTES;
constexpr int SM_REF_BYTES = (RT_SLOTS * RT_STRIDE + RL_SLOTS * RL_STRIDE) * 2;   // 9792
constexpr int SM_MAT = SM_REF + SM_REF_BYTES;
constexpr int SM_MISC = SM_MAT + ((MAT_BYTES + 15) / 16) * 16; // s_dc, work counter, mbarrier
constexpr int SM_DEC = SM_MISC + 32;                            // u32[DEC_MAX]: min over modes of (cost << 6 | mode)
constexpr int SM_TOTAL = SM_DEC + 2048 * 4;

// TMA staging box: frame rows tileY-3 .. tileY+65, columns ctuX-8 .. ctuX+135 (halo 3 >= filter radius 2 + the
// boundary halo 1; 8 columns on the left keep every row 16-byte aligned).  It overlays the s_red scratch.
constexpr int STG_W = 144, STG_H = 69, STG_X0 = 8, STG_Y0 = 3;
constexpr int STG_BYTES = STG_W * STG_H * 2;                   // 19872
static_assert(STG_BYTES <= SM_RED_BYTES, "staging tile must fit into the scratch it overlays");
static_assert(SM_RED % 128 == 0, "TMA destination must be 128-byte aligned");


struct DevType {               // what device code needs to know about a CU type; everything positional is in g_lane
    uint8_t w, h, modes, shape;
    uint8_t parts_log2, narrow, pad[2];   // lanes that share one (CU, mode): 4 for 64x64 (a quarter of the strips each), else 1; narrow: 16-bit entries in the compact table
    uint32_t cost_off, cu_off;       // first cost / first CU of the type inside a CTU
    uint32_t cmp_off;                // byte offset of the type's block inside a CTU's compact record (mip_compact.h)
};

__constant__ DevType c_types[MIP_NUM_TYPES];
// Two splits of a half's work list into chunks are kept side by side: [0] for throughput (frames overlap on the GPU, tails
// are filled by the next frame: few, equal chunks = fewest tile stagings) and [1] for a lone frame (decreasing shares: the
// CTAs that run last are short, so the frame's tail is).  A launch names the one it wants.
__constant__ int c_chunks[2];
__constant__ int c_chunk_begin[2][2][MAX_CHUNKS + 1];
// fused decisions: a chunk never splits the modes of a CU, so the CTA owns the argmin of its CUs
__constant__ uint16_t c_chunk_ord[2][2][MAX_CHUNKS + 1];   // first CU ordinal of each chunk (per split and half)
__constant__ uint16_t c_ord2cu[2][MAX_ORD];       // CU ordinal inside a half -> CU index inside the CTU (0..5379)
__device__ uint2 g_lane[2][MAX_WORK][32];      // per (half, warp task, lane): what the lane does, see the task loop of mip_cost_kernel
// (lane-record layout, the REC_* flags and the CU shapes: mip_work_list.h)
__device__ uint8_t g_mat[MAT_BYTES];           // (coef - 32) as signed bytes, padded layout above

static int g_chunks[2] = {0, 0};

// ------------------------------------------------------------------------------------------
// Device helpers
// ------------------------------------------------------------------------------------------
struct Ctx {
    const int* s_orig;        // [128][OS], holds orig + 1
    // reference samples are addressed by 32-bit shared-memory addresses held in registers (as generic pointers the
    // compiler re-derives the shared window base -- an S2R of SR_CgaCtaId -- three or four times per warp task)
    uint32_t a_refT;          // row slots:    sample (slot, x) at 2 * (slot * RT_STRIDE + 8 + x)
    uint32_t a_refL;          // column slots: sample (slot, y) at 2 * (slot * RL_STRIDE + 8 + y)
    uint32_t a_dc;            // one cell holding 1 << (bitDepth - 1)
    int maxv;                 // (1 << bitDepth) - 1
    uint32_t maxv2;           // maxv in both 16-bit halves
    uint32_t a_red;           // this thread's column of the [RED_WORDS][NT] scratch (shared-memory address)
    const uint8_t* s_mat;
    int ctuX, tileY;          // frame position of the tile origin
    bool topEdge, leftEdge;   // the tile touches the frame's first row / first column (uniform over the CTA)
};

__device__ __forceinline__ int ilog2c(int v) { return v == 1 ? 0 : v == 2 ? 1 : v == 4 ? 2 : v == 8 ? 3 : v == 16 ? 4 : 5; }

// 4x4 SATD of kernel_aux_functions.cl:142-249.  e = (orig - pred) + 1 in raster order: every difference carries the
// same +1 (see diff_*() below), which reaches exactly one Hadamard coefficient, the DC one, as +16.
// Three butterfly stages are explicit; the fourth is folded into the absolute sums:
// |a + b| + |a - b| = |a - (-b)| + |a - b| = two VABSDIFF (abs-diff-accumulate) on (a, -b) and (a, b).
// The DC pair is split because of the mean-scaled DC term (abs(d0) >> 2, :244-245); its -16 rides in the negation.
__device__ __forceinline__ int satd4x4(const int (&e)[16]) {
    int m[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int a0 = e[i] + e[12 + i], a1 = e[4 + i] + e[8 + i];
        int a2 = e[4 + i] - e[8 + i], a3 = e[i] - e[12 + i];
        m[i] = a0 + a1; m[4 + i] = a2 + a3; m[8 + i] = a0 - a1; m[12 + i] = a3 - a2;
    }
    int s = 0, dc = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int b0 = m[4 * r] + m[4 * r + 3], b1 = m[4 * r + 1] + m[4 * r + 2];
        const int b2 = m[4 * r + 1] - m[4 * r + 2], b3 = m[4 * r] - m[4 * r + 3];
        if (r == 0) dc = __sad(b0, 16 - b1, 0);      // |b0 + b1 - 16| = |coefficient 0|
        else s = __sad(b0, -b1, s);
        s = __sad(b0, b1, s);
        s = __sad(b3, -b2, s);
        s = __sad(b3, b2, s);
    }
    return (s + (dc >> 2) + 1) >> 1;
}

// Differences are kept as e = orig - pred + 1, one instruction each whatever the prediction looks like:
//  * the original-sample tile holds o + 1, so against a sample p that needs no shift e = (o + 1) - p;
//  * against an interpolated sample floor(v / 2^s):  o - (v >> s) = (o + 1) + (~v >> s)  (two's complement, one LEA.HI),
//    and e = that + 1 = (o + 1) + ((~v + 2^s) >> s): the 2^s is folded into the start value of the running ~v.
// The +1 is free downstream: SAD takes |e - 1| (VABSDIFF against 1), the SATD corrects its DC coefficient (above).
__device__ __forceinline__ int diff_shifted(int o1, int nv, int s) { return o1 + (nv >> s); }
__device__ __forceinline__ int diff_plain(int o1, int p) { return o1 - p; }
// -(a + b) as one IADD3 with both operands negated; opaque so that LLVM cannot rewrite (-(a + b)) >> 1 into a
// shift plus a subtraction, which would cost the LEA.HI of diff_shifted().
__device__ __forceinline__ int neg_sum(int a, int b) {
    int t;
    asm("{\n\t.reg .s32 u;\n\tadd.s32 u, %1, %2;\n\tneg.s32 %0, u;\n\t}" : "=r"(t) : "r"(a), "r"(b));
    return t;
}
// clamp to the sample range in one VIMNMX.RELU: max(min(v, maxv), 0), maxv = 1023 for the reference's 10 bits (intra.cl:482)
__device__ __forceinline__ int clamp_px(int v, int maxv) { return __vimin_s32_relu(v, maxv); }

// One 4x4 block given its 16 offset differences e = orig - pred + 1 (raster): SAD += sum|e - 1| (one VABSDIFF
// each), SATD += satd4x4(e).
__device__ __forceinline__ void block_cost_d(const int (&e)[16], int& sad, int& satd) {
#pragma unroll
    for (int k = 0; k < 16; ++k) sad = __sad(e[k], 1, sad);
    satd += satd4x4(e);
}

__device__ __forceinline__ void load_o1_row(const int* o, int (&v)[4]) {
    const int4 t = *reinterpret_cast<const int4*>(o);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

// The reduced prediction of a lane is a column of the [RED_WORDS][NT] scratch: sample (j, c) of an R-wide block is half
// (c & 1) of word j * R / 2 + (c >> 1), words NT * 4 bytes apart.  The up-sampling loops walk it with byte pointers so
// that every load has an immediate offset: StripPtr = where a strip's first reduced sample (and the one before it)
// sits in row 0; a row is RED_ROWB<R> bytes further down.
constexpr int RED_WB = NT * 4;                                     // bytes between consecutive words of a lane's column
template <int R> constexpr int RED_ROWB = (R / 2) * RED_WB;        // bytes between reduced rows
// The scratch column is read and written through shared-memory addresses with one LDS.U16 per sample: given pointers the
// compiler fuses the loads of two neighbouring samples into an LDS.32 plus two extractions -- three instructions where two
// do, in a kernel bound by instruction issue.  Volatile: stores and loads of the column keep their program order.
__device__ __forceinline__ int ld_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return (int)v;
}
__device__ __forceinline__ void st_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }

struct StripPtr { uint32_t a, b; bool first; int odd; };
template <int UH>
__device__ __forceinline__ StripPtr strip_ptr(uint32_t base, int s) {
    StripPtr p;
    if constexpr (UH == 1) { p.a = base + 2 * s * RED_WB; p.b = p.a; p.first = false; p.odd = 0; }          // c = 4s .. 4s+3
    else if constexpr (UH == 2) { p.a = base + s * RED_WB; p.first = s == 0; p.b = p.first ? p.a : p.a - RED_WB + 2; p.odd = 0; }   // c = 2s, 2s+1; before: 2s-1
    else {
        const int c = UH == 4 ? s : (s >> 1);                                                                   // one reduced column
        p.a = base + (c >> 1) * RED_WB + (c & 1) * 2;
        p.first = c == 0;
        p.b = (c & 1) ? p.a - 2 : (p.first ? p.a : p.a - RED_WB + 2);   // the first strip takes refL instead: any valid address
        p.odd = s & 1;
    }
    return p;
}

// Horizontal up-sampling (A.4, intra.cl:818-844) of one reduced row for the four columns of a strip.  pa / pb = the
// strip's pointers moved to that row; Lval = refL[y] of the row (the sample before the first reduced column).
__device__ __forceinline__ int avg_up(int a, int b) {   // (a + b + 1) >> 1 as IADD3 + SHF
    int t;
    asm("add.s32 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(b));
    return (t + 1) >> 1;
}

template <int UH>
__device__ __forceinline__ void hor_row(uint32_t pa, uint32_t pb, bool first, int odd, int Lval, int (&cur)[4]) {
    if constexpr (UH == 1) {
        cur[0] = ld_u16(pa); cur[1] = ld_u16(pa + 2); cur[2] = ld_u16(pa + RED_WB); cur[3] = ld_u16(pa + RED_WB + 2);
    } else if constexpr (UH == 2) {
        const int a0 = ld_u16(pa), a1 = ld_u16(pa + 2);
        const int ldb = ld_u16(pb), bef = first ? Lval : ldb;
        cur[0] = avg_up(bef, a0); cur[1] = a0; cur[2] = avg_up(a0, a1); cur[3] = a1;
    } else if constexpr (UH == 4) {
        const int a = ld_u16(pa);
        const int ldb = ld_u16(pb), bef = first ? Lval : ldb;
        const int dl = a - bef, v = 4 * bef + 2;
        cur[0] = (v + dl) >> 2; cur[1] = (v + 2 * dl) >> 2; cur[2] = (v + 3 * dl) >> 2; cur[3] = a;
    } else {  // UH == 8: two strips per reduced column
        const int a = ld_u16(pa);
        const int ldb = ld_u16(pb), bef = first ? Lval : ldb;
        const int dl = a - bef, v = 8 * bef + 4 + odd * 4 * dl;
        cur[0] = (v + dl) >> 3; cur[1] = (v + 2 * dl) >> 3; cur[2] = (v + 3 * dl) >> 3; cur[3] = (v + 4 * dl) >> 3;
    }
}

// Reduced boundary (intra.cl:127-141, 260-279): B rounded means over D consecutive 16-bit samples each, p 8-byte aligned
// (CU positions are multiples of 4).  64-bit shared loads; IDP.2A against (1, 1) adds a pair of samples per instruction
// and starts from the rounding offset D / 2.
__device__ __forceinline__ int lds_u16(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return (int)v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

template <int B, int D>
__device__ __forceinline__ void reduce_bdry(uint32_t p, int (&red)[B]) {
    if constexpr (D == 1) {
        const uint2 v = lds_v2(p);
        red[0] = v.x & 0xffff; red[1] = v.x >> 16; red[2] = v.y & 0xffff; red[3] = v.y >> 16;
    } else {
        uint32_t w[B * D / 2];
#pragma unroll
        for (int i = 0; i < B * D / 4; ++i) {
            const uint2 v = lds_v2(p + 8 * i);
            w[2 * i] = v.x; w[2 * i + 1] = v.y;
        }
#pragma unroll
        for (int q = 0; q < B; ++q) {
            unsigned acc = D >> 1;
#pragma unroll
            for (int t = 0; t < D / 2; ++t) acc = __dp2a_lo(w[q * (D / 2) + t], 0x0101u, acc);
            red[q] = (int)(acc >> ilog2c(D));
        }
    }
}

// One (CU, mode): everything from the boundaries to SAD/SATD.  SID = sizeId, W x H = CU size.
// cuX, cuY are relative to the tile.  PARTS lanes share the (CU, mode): lane `part` takes a contiguous
// 1/PARTS of the strips (every lane still computes the whole reduced prediction into its own scratch).
template <int SID, int W, int H, int PARTS = 1>
__device__ __forceinline__ void run_task(const Ctx& c, int cuX, int cuY, int mode, int part, int& sad, int& satd) {
    constexpr int R = SID == 2 ? 8 : 4;                   // reduced prediction side
    constexpr int B = SID == 0 ? 2 : 4;                   // reduced boundary samples per side
    constexpr int M = SID == 2 ? 6 : (SID == 1 ? 8 : 16); // matrices
    constexpr int UH = W / R, UV = H / R;

    // ---- A.1 complete boundaries: pointer + stride into the reference tile (intra.cl:96-107, 232-243)
    // Inside the frame: row Y-1 is row slot cuY/4, column X-1 is column slot cuX/4.  On the frame's first row / column the
    // boundary is one replicated sample -- F[0][X-1] resp. F[Y-1][0], or the mid-grey cell in the frame's corner -- read with
    // stride 0; whether the tile touches those edges at all is uniform over the CTA, so everywhere else the test is one
    // uniform branch.
    uint32_t T = c.a_refT + 2 * ((cuY >> 2) * RT_STRIDE + 8 + cuX), L = c.a_refL + 2 * ((cuX >> 2) * RL_STRIDE + 8 + cuY);
    int stT = 2, stL = 2;      // byte stride between samples: 2, or 0 for a replicated sample
    if (c.topEdge || c.leftEdge) {
        const bool atTop = c.topEdge && cuY == 0, atLeft = c.leftEdge && cuX == 0;
        if (atTop) { T = atLeft ? c.a_dc : c.a_refT + 2 * (RT_ROW0 * RT_STRIDE + 8 + cuX - 1); stT = 0; }
        if (atLeft) { L = atTop ? c.a_dc : c.a_refL + 2 * (RL_COL0 * RL_STRIDE + 8 + cuY - 1); stL = 0; }
    }

    // ---- A.2 reduced boundaries (intra.cl:127-141, 260-279)
    constexpr int DT = W / B, DL = H / B;
    int bd[2 * B];
    {
        int redT[B], redL[B];
        if (stT) reduce_bdry<B, DT>(T, redT);
        else {   // frame edge: one replicated sample v, and (D * v + D / 2) >> log2(D) == v
            const int v = lds_u16(T);
#pragma unroll
            for (int q = 0; q < B; ++q) redT[q] = v;
        }
        if (stL) reduce_bdry<B, DL>(L, redL);
        else {
            const int v = lds_u16(L);
#pragma unroll
            for (int q = 0; q < B; ++q) redL[q] = v;
        }
        const bool tr = mode >= M;
#pragma unroll
        for (int i = 0; i < B; ++i) { bd[i] = tr ? redL[i] : redT[i]; bd[B + i] = tr ? redT[i] : redL[i]; }
    }
    // ---- A.3 input vector, packed s16x2 for IDP.2A (intra.cl:434-452)
    const int first = bd[0];
    // ((32 + sum) >> 6) + first == (32 + 64 * first + sum) >> 6   (intra.cl:454, 481).  For sizeId 1 and 2 everything is
    // carried 4x (inputs and start value): (4 * acc) >> 8 == acc >> 6, and a shift by 8 is a byte selection, so one PRMT
    // both shifts and packs two samples and one VIMNMX.S16x2.RELU clamps both (the 16-bit range holds: kernels_init checks
    // the matrices' row sums).
    constexpr int SC = SID == 0 ? 0 : 2;
    const int acc0 = (32 + 64 * first) << SC;
    int ipk[B];
    {
        int in[2 * B];
        const int fs = first << SC;
        in[0] = (SID == 2) ? 0 : (((c.maxv + 1) >> 1) << SC) - fs;
#pragma unroll
        for (int i = 1; i < 2 * B; ++i) in[i] = (bd[i] << SC) - fs;        // one shift-and-add each
#pragma unroll
        for (int k = 0; k < B; ++k) ipk[k] = __byte_perm(in[2 * k], in[2 * k + 1], 0x5410);   // the low halves of two inputs
    }
    // the transposed copies of a size's matrices follow the plain ones at the same pitch (static_assert below), and a
    // transposed mode is M + the matrix number: both kinds sit at offset + mode * pitch

    if constexpr (SID == 0) {
        // 4x4: the reduced prediction is the prediction (intra.cl:726-727, 934-935)
        const uint8_t* mb = c.s_mat + M0_OFF + mode * M0_STRIDE;   // row = output position, also for transposed modes
        int p[16];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int4 cw4 = *reinterpret_cast<const int4*>(mb + a * 16);      // the 4 taps of output samples 4a .. 4a+3
            const int cw[4] = {cw4.x, cw4.y, cw4.z, cw4.w};
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int acc = acc0;
                acc = __dp2a_lo(ipk[0], cw[b], acc);
                acc = __dp2a_hi(ipk[1], cw[b], acc);
                p[a * 4 + b] = clamp_px(acc >> 6, c.maxv);
            }
        }
        int d[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int o1[4];
            load_o1_row(c.s_orig + (cuY + i) * OS + cuX, o1);
#pragma unroll
            for (int k = 0; k < 4; ++k) d[4 * i + k] = diff_plain(o1[k], p[4 * i + k]);
        }
        block_cost_d(d, sad, satd);
        return;
    } else {
        // ---- A.3 matrix-vector product -> private scratch column, two samples per word
        const uint8_t* mb = c.s_mat + (SID == 2 ? M2_OFF + mode * M2_STRIDE : M1_OFF + mode * M1_STRIDE);
        const char* mp = reinterpret_cast<const char*>(mb);     // matrix rows a*R + b (8 taps each), two per 16-byte load
        uint32_t wp = c.a_red;                                  // words of this lane's scratch column, RED_WB bytes apart
        // rolled, because the kernel lives on the instruction caches' good will; two rows per trip for the 4x4 reduced
        // prediction, whose row is only two stores long (measured: 1 % of the frame; unrolling further, or the block loops, costs more than it saves)
#pragma unroll (R == 4 ? 2 : 1)
        for (int a = 0; a < R; ++a) {
#pragma unroll
            for (int b = 0; b < R; b += 2) {
                const int4 cw = *reinterpret_cast<const int4*>(mp + b * 8);
                int acc = acc0, acd = acc0;
                acc = __dp2a_lo(ipk[0], cw.x, acc);  acd = __dp2a_lo(ipk[0], cw.z, acd);
                acc = __dp2a_hi(ipk[1], cw.x, acc);  acd = __dp2a_hi(ipk[1], cw.z, acd);
                acc = __dp2a_lo(ipk[2], cw.y, acc);  acd = __dp2a_lo(ipk[2], cw.w, acd);
                acc = __dp2a_hi(ipk[3], cw.y, acc);  acd = __dp2a_hi(ipk[3], cw.w, acd);
                // bytes 1..2 of each accumulator = (4 * acc) >> 8 as a signed 16-bit value; clamp both halves to 0..maxv at once
                st_u32(wp + (b >> 1) * RED_WB, __vimin_s16x2_relu(__byte_perm(acc, acd, 0x6521), c.maxv2));
            }
            mp += R * 8;
            wp += RED_ROWB<R>;
        }
        // ---- A.4 + A.5 strip-wise: 4 columns at a time, top to bottom
        const int* orig = c.s_orig + cuY * OS + cuX;
        constexpr int STRIPS = W / 4 / PARTS;
#pragma unroll 1
        for (int s = part * STRIPS; s < (part + 1) * STRIPS; ++s) {
            const int x0 = 4 * s;
            const StripPtr sp = strip_ptr<UH>(c.a_red, s);
            constexpr int ROWB = RED_ROWB<R>;
            int prev[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) prev[k] = lds_u16(T + (x0 + k) * stT);
            if constexpr (UV >= 4) {
                constexpr int SH = UV == 4 ? 2 : 3;
#pragma unroll 1
                for (int j = 0; j < R; ++j) {
                    int cur[4];
                    hor_row<UH>(sp.a + j * ROWB, sp.b + j * ROWB, sp.first, sp.odd, lds_u16(L + (j * UV + UV - 1) * stL), cur);
                    // nv = ~(UV*prev + UV/2 + i*dl) + UV walks down the rows; see diff_shifted()
                    int dl[4], nv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { dl[k] = prev[k] - cur[k]; nv[k] = -UV * prev[k] + ((UV >> 1) - 1); }
#pragma unroll
                    for (int blk = 0; blk < UV / 4; ++blk) {
                        int d[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            int o1[4];
                            load_o1_row(orig + (j * UV + blk * 4 + i) * OS + x0, o1);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (blk * 4 + i == UV - 1) d[4 * i + k] = diff_plain(o1[k], cur[k]);
                                else { nv[k] += dl[k]; d[4 * i + k] = diff_shifted(o1[k], nv[k], SH); }
                            }
                        }
                        block_cost_d(d, sad, satd);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) prev[k] = cur[k];
                }
            } else if constexpr (UV == 2) {
#pragma unroll 1
                for (int jp = 0; jp < R / 2; ++jp) {
                    int c0[4], c1[4], d[16], o1[4];
                    hor_row<UH>(sp.a + 2 * jp * ROWB, sp.b + 2 * jp * ROWB, sp.first, sp.odd, lds_u16(L + (4 * jp + 1) * stL), c0);
                    hor_row<UH>(sp.a + (2 * jp + 1) * ROWB, sp.b + (2 * jp + 1) * ROWB, sp.first, sp.odd, lds_u16(L + (4 * jp + 3) * stL), c1);
                    const int* o = orig + (4 * jp) * OS + x0;
                    load_o1_row(o, o1);            // row 0: (prev + c0 + 1) >> 1
#pragma unroll
                    for (int k = 0; k < 4; ++k) d[k] = diff_shifted(o1[k], neg_sum(prev[k], c0[k]), 1);     // ~(prev + c0 + 1) + 2
                    load_o1_row(o + OS, o1);       // row 1: c0
#pragma unroll
                    for (int k = 0; k < 4; ++k) d[4 + k] = diff_plain(o1[k], c0[k]);
                    load_o1_row(o + 2 * OS, o1);   // row 2: (c0 + c1 + 1) >> 1
#pragma unroll
                    for (int k = 0; k < 4; ++k) d[8 + k] = diff_shifted(o1[k], neg_sum(c0[k], c1[k]), 1);
                    load_o1_row(o + 3 * OS, o1);   // row 3: c1
#pragma unroll
                    for (int k = 0; k < 4; ++k) { d[12 + k] = diff_plain(o1[k], c1[k]); prev[k] = c1[k]; }
                    block_cost_d(d, sad, satd);
                }
            } else {  // UV == 1: reduced rows are pixel rows
#pragma unroll 1
                for (int jq = 0; jq < R / 4; ++jq) {
                    int d[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        int cur[4], o1[4];
                        hor_row<UH>(sp.a + (4 * jq + i) * ROWB, sp.b + (4 * jq + i) * ROWB, sp.first, sp.odd, lds_u16(L + (4 * jq + i) * stL), cur);
                        load_o1_row(orig + (4 * jq + i) * OS + x0, o1);
#pragma unroll
                        for (int k = 0; k < 4; ++k) d[4 * i + k] = diff_plain(o1[k], cur[k]);
                    }
                    block_cost_d(d, sad, satd);
                }
            }
        }
    }
}

// One warp task of a known shape: the in-frame test with the CU size as immediates, the work, and -- for the 64x64 type,
// whose (CU, mode) pairs take four lanes -- the sum over the four strip groups.  Returns whether this lane's CU is inside the frame.
template <int SID, int W, int H, int PARTS = 1>
__device__ __forceinline__ bool do_task(const Ctx& c, int cuX, int cuY, int mode, int part, bool inRange, int rowsValid, int frameW, int& sad, int& satd) {
    const bool active = inRange && cuY + H <= rowsValid && c.ctuX + cuX + W <= frameW;   // CU fully inside the frame
    if (__any_sync(0xffffffffu, active)) {
        run_task<SID, W, H, PARTS>(c, cuX, cuY, mode, part, sad, satd);
        if constexpr (PARTS == 4) {   // lanes 4k..4k+3 hold the four strip groups of one (CU, mode)
            sad += __shfl_xor_sync(0xffffffffu, sad, 1);  satd += __shfl_xor_sync(0xffffffffu, satd, 1);
            sad += __shfl_xor_sync(0xffffffffu, sad, 2);  satd += __shfl_xor_sync(0xffffffffu, satd, 2);
        }
    }
    return active;
}

// ------------------------------------------------------------------------------------------
// TMA + mbarrier (sm_90+/sm_100a PTX)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// one 2-D box of the frame -> shared memory; completion (bytes) is counted on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

// ------------------------------------------------------------------------------------------
// Low-pass filter of one sample taken from the staged tile (A.6; same arithmetic as mip_filter_kernel below).
// (x, y) = frame position; stg points at the staged sample of frame position (x, y); taps outside the frame are
// excluded by coordinates (the TMA zero fill is never interpreted as a sample).
// ------------------------------------------------------------------------------------------
template <int RAD, bool IS2D>
__device__ __forceinline__ int filter_staged(const uint16_t* stg, int x, int y, int W, int H, int kidx) {
    int num = 0, den = 0;
    if constexpr (IS2D) {
#pragma unroll
        for (int dy = -RAD; dy <= RAD; ++dy)
#pragma unroll
            for (int dx = -RAD; dx <= RAD; ++dx) {
                const int xx = x + dx, yy = y + dy;
                if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                const int k = RAD == 1 ? mip_k3(kidx, dy, dx) : mip_k5(kidx, dy, dx);
                num += k * (int)stg[dy * STG_W + dx];
                den += k;
            }
    } else {
        int k[2 * RAD + 1];
#pragma unroll
        for (int i = 0; i <= 2 * RAD; ++i) k[i] = RAD == 1 ? mip_k3(kidx, -1, i - 1) : mip_k5(kidx, -2, i - 2);
#pragma unroll
        for (int dy = -RAD; dy <= RAD; ++dy) {
            int hor = 0;
#pragma unroll
            for (int dx = -RAD; dx <= RAD; ++dx) {
                const int xx = x + dx, yy = y + dy;
                if (xx >= 0 && xx < W && yy >= 0 && yy < H) hor += k[dx + RAD] * (int)stg[dy * STG_W + dx];
            }
            num += k[dy + RAD] * hor;
        }
        if constexpr (RAD == 1) {
            const int nEdges = (x == 0) + (x == W - 1) + (y == 0) + (y == H - 1);
            const int k0 = k[0], k1 = k[1];
            den = nEdges >= 2 ? (k0 + 2 * k1 + k1 * k1) : (nEdges == 1 ? (2 * k0 + 3 * k1 + k1 * k1) : (4 * k0 + 4 * k1 + k1 * k1));
        } else {
            auto ksum = [&](int i0, int j0) {
                int t = 0;
                for (int i = i0; i < 5; ++i)
                    for (int j = j0; j < 5; ++j) t += mip_k5(kidx, i - 2, j - 2);
                return t;
            };
            const bool oTB = (y == 0) || (y == H - 1), iTB = (y == 1) || (y == H - 2);
            const bool oLR = (x == 0) || (x == W - 1), iLR = (x == 1) || (x == W - 2);
            const bool oC = oTB && oLR, iC = iTB && iLR;
            const bool ifc = (oLR && iTB) || (iLR && oTB);
            const bool oE = !oC && !ifc && (oTB || oLR), iE = !iC && !ifc && (iTB || iLR);
            den = ksum(0, 0);
#pragma unroll
            for (int dy = -2; dy <= 2; ++dy)
                if (y + dy < 0 || y + dy >= H) den -= k[dy + 2];
            if (oC) den = ksum(2, 2);
            if (iC) den = ksum(1, 1);
            if (oE) den = ksum(0, 2);
            if (iE) den = ksum(0, 1);
            if (ifc) den = ksum(1, 2);
        }
    }
    return (num + den / 2) / den;
}

// Reference tile from the staged tile: a copy (FT == 0, original samples) or the low-pass filtered samples.
// Item i < RT_SLOTS*129 is sample cc (frame column ctuX-1+cc) of row slot i/129; the rest are the column slots.
// Frame position -> (slot): rows tileY-1, tileY+3, .., tileY+63 and frame row 0 (slot RT_ROW0, only when tileY == 0);
// columns ctuX-1, ctuX+3, .., ctuX+127 and frame column 0 (slot RL_COL0, only when ctuX == 0).
template <int RAD, bool IS2D, bool FILTER>
__device__ __forceinline__ void build_ref_tile(uint16_t* s_refT, uint16_t* s_refL, const uint16_t* stg, int ctuX, int tileY,
                                               int W, int H, const FilterParams& fp, int tid) {
    constexpr int N = 2 * RAD + 1;
    constexpr int NROW = RT_SLOTS * 129, NCOL = RL_SLOTS * 65;
    for (int i = tid; i < NROW + NCOL; i += NT) {
        int ty, tx;             // tile-relative position of the sample: ty in -1..63, tx in -1..127
        uint16_t* dst;
        if (i < NROW) {
            const int slot = i / 129, cc = i - slot * 129;
            ty = slot == RT_ROW0 ? 0 : 4 * slot - 1;
            tx = cc - 1;
            dst = s_refT + slot * RT_STRIDE + 8 + tx;
        } else {
            // slot fastest: neighbouring lanes read the staged tile 4 samples apart in the same row (2-way bank
            // conflicts); row fastest would walk down a column of the 72-word-pitch staging tile (8-way)
            const int k = i - NROW, rr = k / RL_SLOTS, slot = k - rr * RL_SLOTS;
            tx = slot == RL_COL0 ? 0 : 4 * slot - 1;
            ty = rr - 1;
            dst = s_refL + slot * RL_STRIDE + 8 + ty;
        }
        const int y = tileY + ty, x = ctuX + tx;
        int v = 0;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const uint16_t* p = stg + (ty + STG_Y0) * STG_W + (tx + STG_X0);
            if constexpr (!FILTER) {
                v = *p;
            } else if (x >= RAD && x < W - RAD && y >= RAD && y < H - RAD) {
                // whole window inside the frame: fixed weights, fixed denominator, division by multiply-high
                int num = fp.full_den >> 1;
#pragma unroll
                for (int dy = -RAD; dy <= RAD; ++dy)
#pragma unroll
                    for (int dx = -RAD; dx <= RAD; ++dx) num += fp.coef[(dy + RAD) * N + dx + RAD] * (int)p[dy * STG_W + dx];
                v = (int)__umulhi((uint32_t)num, fp.full_magic);
            } else {
                v = filter_staged<RAD, IS2D>(p, x, y, W, H, fp.kidx);      // frame border: position classes of A.6
            }
        }
        *dst = (uint16_t)v;
    }
}

// ------------------------------------------------------------------------------------------
// The fused kernel
// ------------------------------------------------------------------------------------------
// COMPACT: the cost table is the compact one (mip_compact.h).  A template parameter, not an argument: the int32 path pays
// nothing for the other's existence (as an argument it cost 0.6 % of the frame).
// OUT: which results the launch writes -- template parameters, not pointer tests per warp task (a frame is 390 000 tasks).
constexpr int OUT_COST = 1, OUT_SADSATD = 2, OUT_DEC = 4;
template <bool COMPACT, int OUT>
__global__ void __launch_bounds__(NT, 2)
mip_cost_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FilterParams fp, int W, int H, int split, int maxv,
                int32_t* __restrict__ g_cost, int32_t* __restrict__ g_sad, int32_t* __restrict__ g_satd,
                uint8_t* __restrict__ g_best_mode, int32_t* __restrict__ g_best_cost) {
    extern __shared__ __align__(128) unsigned char smem[];
    int* s_orig = reinterpret_cast<int*>(smem + SM_ORIG);
    uint16_t* s_refT = reinterpret_cast<uint16_t*>(smem + SM_REF);
    uint16_t* s_refL = s_refT + RT_SLOTS * RT_STRIDE;
    uint32_t* s_red = reinterpret_cast<uint32_t*>(smem + SM_RED);
    uint16_t* s_stg = reinterpret_cast<uint16_t*>(smem + SM_RED);    // TMA box, dead before the first task starts
    uint8_t* s_mat = smem + SM_MAT;
    uint16_t* s_dc = reinterpret_cast<uint16_t*>(smem + SM_MISC);
    int* s_next = reinterpret_cast<int*>(smem + SM_MISC + 4);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + SM_MISC + 16);
    uint32_t* s_dec = reinterpret_cast<uint32_t*>(smem + SM_DEC);

    const int tid = threadIdx.x, lane = tid & 31;
    // chunk-major unit order: CTAs that are resident together work on the same chunk (the same CU
    // shapes, hence the same code) of different CTU halves, which keeps the instruction caches warm
    const int ctuCols = (W + 127) >> 7;
    const int halves = 2 * ctuCols * ((H + 127) >> 7);
    const int chunk = blockIdx.x / halves, hu = blockIdx.x - chunk * halves;
    const int ctu = hu >> 1, half = hu & 1;
    const int ctuX = (ctu % ctuCols) << 7, tileY = ((ctu / ctuCols) << 7) + half * TILE_ROWS;
    const int rowsValid = min(TILE_ROWS, H - tileY);     // <= 0: the whole half lies below the frame
    const int wbeg = c_chunk_begin[split][half][chunk], wcnt = c_chunk_begin[split][half][chunk + 1] - wbeg;
    const uint32_t ctuBase = (uint32_t)ctu * MIP_COSTS_PER_CTU;   // 32-bit element index: frames up to 43 000 CTUs

    if (rowsValid > 0) {
        // ---- stage: one TMA box (tile + halo) signalled on an mbarrier; the matrices come in meanwhile
        if (tid == 0) {
            mbar_init(s_bar, 1);
            mbar_expect_tx(s_bar, STG_BYTES);
            tma_load_2d(s_stg, &tmap, ctuX - STG_X0, tileY - STG_Y0, s_bar);
        }
        for (int i = tid; i < MAT_BYTES / 4; i += NT)
            reinterpret_cast<uint32_t*>(s_mat)[i] = reinterpret_cast<const uint32_t*>(g_mat)[i];
        __syncthreads();                                  // the barrier word is initialised for everyone
        if (!mbar_try_wait(s_bar, 0)) {
            // a lost TMA must not hang the GPU: give up after 4 s of wall time (time-based, so that a sanitizer or a
            // debugger slowing the kernel down by orders of magnitude cannot trip it)
            const uint64_t t0 = globaltimer_ns();
            while (!mbar_try_wait(s_bar, 0))
                if (globaltimer_ns() - t0 > 4000000000ull) __trap();
        }
        // ---- originals (+1) as int32 (see diff_shifted()); rows below the frame are TMA zero fill
        for (int i = tid; i < TILE_ROWS * 16; i += NT) {
            const int y = i >> 4, xc = (i & 15) << 3;
            const uint4 v = *reinterpret_cast<const uint4*>(s_stg + (y + STG_Y0) * STG_W + STG_X0 + xc);
            int4* dst = reinterpret_cast<int4*>(s_orig + y * OS + xc);
            dst[0] = make_int4((v.x & 0xffff) + 1, (v.x >> 16) + 1, (v.y & 0xffff) + 1, (v.y >> 16) + 1);
            dst[1] = make_int4((v.z & 0xffff) + 1, (v.z >> 16) + 1, (v.w & 0xffff) + 1, (v.w >> 16) + 1);
        }
        // ---- reference row/column slots: copy, or the fused low-pass filter (fp.type = 1..8)
        switch (fp.type) {
            case 0:         build_ref_tile<1, true, false>(s_refT, s_refL, s_stg, ctuX, tileY, W, H, fp, tid); break;
            case 1: case 2: build_ref_tile<1, false, true>(s_refT, s_refL, s_stg, ctuX, tileY, W, H, fp, tid); break;
            case 3: case 4: build_ref_tile<1, true, true>(s_refT, s_refL, s_stg, ctuX, tileY, W, H, fp, tid); break;
            case 5: case 6: build_ref_tile<2, false, true>(s_refT, s_refL, s_stg, ctuX, tileY, W, H, fp, tid); break;
            default:        build_ref_tile<2, true, true>(s_refT, s_refL, s_stg, ctuX, tileY, W, H, fp, tid); break;
        }
    }
    const int ordBeg = c_chunk_ord[split][half][chunk], ordCnt = c_chunk_ord[split][half][chunk + 1] - ordBeg;
    if constexpr (OUT & OUT_DEC)
        for (int i = tid; i < ordCnt; i += NT) s_dec[i] = 0xffffffffu;
    if (tid == 0) { *s_dc = (uint16_t)((maxv + 1) >> 1); *s_next = 0; }
    __syncthreads();                                      // tiles complete; the staging box may now be overwritten

    Ctx c;
    c.s_orig = s_orig;
    asm volatile("mov.u32 %0, %1;" : "=r"(c.a_refT) : "r"(smem_u32(s_refT)));
    asm volatile("mov.u32 %0, %1;" : "=r"(c.a_refL) : "r"(smem_u32(s_refL)));
    asm volatile("mov.u32 %0, %1;" : "=r"(c.a_dc) : "r"(smem_u32(s_dc)));
    c.maxv = maxv;
    c.maxv2 = (uint32_t)maxv * 0x10001u;
    asm volatile("mov.u32 %0, %1;" : "=r"(c.a_red) : "r"(smem_u32(s_red + tid)));
    c.s_mat = s_mat;
    c.ctuX = ctuX;
    c.tileY = tileY;
    c.topEdge = tileY == 0;
    c.leftEdge = ctuX == 0;

    // Warp tasks are drawn from a shared-memory counter, one ahead: the record of the next task (one coalesced 8-byte load
    // from a table built once on the host -- the same 870 KB for every CTU, so it lives in L2) is requested before the
    // current task's arithmetic starts, which hides the atomic + L2 latency of the draw behind ~1000 instructions of work.
    // Everything a task needs is in its record (layout: see g_lane): no table look-up per task.
    // Shared-memory addresses used once per task are kept in registers (left to itself the compiler re-derives them from
    // SR_CgaCtaId -- an S2R -- in every iteration).
    // Lane 0 draws.  ptxas wraps an atomic on a warp-uniform address into a vote / popc / shuffle aggregation (17
    // instructions); adding bits of a loaded record that are always 0 -- which the compiler cannot know -- makes the address
    // formally per-lane and leaves a single predicated ATOMS.ADD.
    const uint2 lr0 = __ldg(&g_lane[half][min(wbeg, MAX_WORK - 2)][lane]);
    uint32_t a_next, a_dec;
    asm volatile("mov.u32 %0, %1;" : "=r"(a_next) : "r"(smem_u32(s_next) + ((lr0.x >> 5) & 4u)));
    asm volatile("mov.u32 %0, %1;" : "=r"(a_dec) : "r"(smem_u32(s_dec) - 4u * (uint32_t)ordBeg));
    // the lane's record of the chunk's first task; past the chunk's end a draw reads the half's last row, which is the end mark
    const char* lane_rec;      // a row of 32 records = 256 bytes; (opaque: the compiler would keep the sum's two terms apart and add them per task)
    asm volatile("mov.u64 %0, %1;" : "=l"(lane_rec) : "l"(reinterpret_cast<const char*>(&g_lane[half][0][lane]) + (size_t)wbeg * 256));
    const int end_row = MAX_WORK - 1 - wbeg;
    auto draw = [&]() -> uint2 {
        int wi;     // only lane 0's value is used
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, 0;\n\t@p atom.shared.add.u32 %0, [%2], 1;\n\t}"
                     : "=r"(wi) : "r"(lane), "r"(a_next) : "memory");
        wi = __shfl_sync(0xffffffffu, wi, 0);
        return __ldg(reinterpret_cast<const uint2*>(lane_rec + (ptrdiff_t)(wi < wcnt ? wi : end_row) * 256));
    };
    uint2 lr = draw();
    while (lr.x != 0xffffffffu) {
        const uint2 lr_next = draw();
        // warp task = 32 consecutive (CU, mode) pairs of one type: a warp touches at most 3-4 CUs, so the shared-memory
        // reads of originals and boundaries are mostly broadcasts (8 CUs x 4 modes per warp measured 60 % more bank conflicts)
        const uint32_t x = lr.x;
        const int cuX = x & 0xff, cuY = __byte_perm(x, 0, 0x4441), mode = __byte_perm(x, 0, 0x4442);
        const bool inRange = (x & REC_INRANGE) != 0;
        const bool writer = (x & REC_WRITER) != 0;
        const uint32_t coff = lr.y & 0x1ffffu;
        const uint32_t sub = lr.y >> 29;
        int sad = 0, satd = 0;
        bool active;
        // Decision = argmin over the CU's modes, lowest mode wins ties: min of (cost << 6 | mode), cost < 2^24, taken in
        // registers (REDUX.MIN) and kept per CU in the CTA's shared-memory table.  How the lanes of a task map to CUs
        // depends on the number of modes, so each dispatch group ends with its own form; lanes that do not vote
        // (outside the frame, past the type's end) enter as 0xffffffff, the table's initial value.
        auto dec_key = [&](bool votes) -> uint32_t { return votes ? (((uint32_t)min(2 * sad, satd) << 6) | (uint32_t)mode) : 0xffffffffu; };
        auto dec_addr = [&]() -> uint32_t { return a_dec + 4u * ((lr.y >> 17) & 0xfffu); };
        auto decide32 = [&]() {      // 32 modes: the warp is one CU (no lane past the type's end), its slot has a single writer
            if constexpr (OUT & OUT_DEC) {
                const uint32_t best = __reduce_min_sync(0xffffffffu, dec_key(active));
                if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(dec_addr()), "r"(best) : "memory");
            }
        };
        auto decide16 = [&]() {      // 16 modes: each half warp is one CU (or lies past the type's end as a whole)
            if constexpr (OUT & OUT_DEC) {
                const uint32_t key = dec_key(writer && active);
                const bool hi = (lane & 16) != 0;
                const uint32_t b0 = __reduce_min_sync(0xffffffffu, hi ? 0xffffffffu : key);
                const uint32_t b1 = __reduce_min_sync(0xffffffffu, hi ? key : 0xffffffffu);
                if (writer && (lane & 15) == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(dec_addr()), "r"(hi ? b1 : b0) : "memory");
            }
        };
        auto decide12 = [&]() {      // 12 modes: a task spans 3-4 CUs and a CU two tasks -- MATCH.ANY finds the lanes of a CU, one of them does an atomicMin
            if constexpr (OUT & OUT_DEC) {
                const bool vote = writer && active;
                const uint32_t key = dec_key(vote);
                const int slot = (int)((lr.y >> 17) & 0xfffu);
                const unsigned grp = __match_any_sync(0xffffffffu, vote ? slot : -1 - lane);
                if (vote) {
                    const uint32_t best = __reduce_min_sync(grp, key);
                    if (lane == __ffs(grp) - 1) asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(dec_addr()), "r"(best) : "memory");
                }
            }
        };
        if ((int)x < 0) {
            active = do_task<0, 4, 4>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd);
            decide32();
        } else if (x & REC_GA32) {
            if (sub == 0) active = do_task<1, 8, 4>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd);
            else active = do_task<1, 4, 8>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd);
            decide16();
        } else if (x & REC_GS1) {
            switch (sub) {
                case 0:  active = do_task<1, 8, 8>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                case 1:  active = do_task<1, 16, 4>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                case 2:  active = do_task<1, 4, 16>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                case 3:  active = do_task<1, 32, 4>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                default: active = do_task<1, 4, 32>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
            }
            decide16();
        } else {
            if (x & REC_G64) active = do_task<2, 64, 64, 4>(c, cuX, cuY, mode, (x >> 24) & 3, inRange, rowsValid, W, sad, satd);
            else switch (sub) {
                case 0:  active = do_task<2, 32, 32>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                case 1:  active = do_task<2, 32, 16>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                case 2:  active = do_task<2, 16, 32>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                case 3:  active = do_task<2, 32, 8>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                case 4:  active = do_task<2, 8, 32>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                case 5:  active = do_task<2, 16, 16>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                case 6:  active = do_task<2, 16, 8>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
                default: active = do_task<2, 8, 16>(c, cuX, cuY, mode, 0, inRange, rowsValid, W, sad, satd); break;
            }
            decide12();
        }
        if (writer) {
            const uint32_t o = ctuBase + coff;
            const int cost = min(2 * sad, satd);            // intra.cl:1166
            if constexpr (OUT & OUT_COST) {
                if constexpr (!COMPACT) g_cost[o] = active ? cost : -1;
                else {
                    // compact table (mip_compact.h): the type's block inside the CTU's record, 16-bit entries for CUs of at
                    // most 32 samples (cost <= 65 472 with 10-bit samples), int32 otherwise
                    int lo = 0, hi = MIP_NUM_TYPES - 1;      // the type whose block holds this cost (last type with cost_off <= coff)
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (c_types[mid].cost_off <= coff) lo = mid; else hi = mid - 1;
                    }
                    const DevType& ty = c_types[lo];
                    unsigned char* rec = reinterpret_cast<unsigned char*>(g_cost) + (size_t)ctu * MIP_COMPACT_BYTES_PER_CTU + ty.cmp_off;
                    const uint32_t idx = coff - ty.cost_off;
                    if (ty.narrow) reinterpret_cast<uint16_t*>(rec)[idx] = active ? (uint16_t)cost : (uint16_t)0xFFFFu;
                    else reinterpret_cast<int32_t*>(rec)[idx] = active ? cost : -1;
                }
            }
            if constexpr (OUT & OUT_SADSATD) { g_sad[o] = active ? sad : -1; g_satd[o] = active ? satd : -1; }   // both or neither (launch_costs)
        }
        lr = lr_next;
    }
    if constexpr (OUT & OUT_DEC) {
        __syncthreads();
        const size_t cuBase = (size_t)ctu * MIP_CUS_PER_CTU;
        for (int i = tid; i < ordCnt; i += NT) {
            const uint32_t v = s_dec[i];
            const size_t o = cuBase + c_ord2cu[half][ordBeg + i];
            g_best_mode[o] = v == 0xffffffffu ? (uint8_t)0xFF : (uint8_t)(v & 63u);
            g_best_cost[o] = v == 0xffffffffu ? -1 : (int32_t)(v >> 6);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Low-pass filters for alternative samples (A.6; intra.cl:1639-3823).  One thread per pixel.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int px_or_zero(const uint16_t* f, int W, int H, int x, int y) {
    return (x >= 0 && x < W && y >= 0 && y < H) ? (int)__ldg(f + (size_t)y * W + x) : 0;
}

template <int RAD, bool IS2D>
__global__ void __launch_bounds__(256)
mip_filter_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int W, int H, int kidx) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    int num = 0, den = 0;
    if constexpr (IS2D) {
        // numerator and denominator over the in-frame taps (intra.cl:2993-3011, 3215-3235)
#pragma unroll
        for (int dy = -RAD; dy <= RAD; ++dy)
#pragma unroll
            for (int dx = -RAD; dx <= RAD; ++dx) {
                const int xx = x + dx, yy = y + dy;
                if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                const int k = RAD == 1 ? mip_k3(kidx, dy, dx) : mip_k5(kidx, dy, dx);
                num += k * (int)__ldg(in + (size_t)yy * W + xx);
                den += k;
            }
    } else {
        // separable passes with the first row of the table; denominator by position class
        int k[2 * RAD + 1];
#pragma unroll
        for (int i = 0; i <= 2 * RAD; ++i) k[i] = RAD == 1 ? mip_k3(kidx, -1, i - 1) : mip_k5(kidx, -2, i - 2);
#pragma unroll
        for (int dy = -RAD; dy <= RAD; ++dy) {
            int hor = 0;
#pragma unroll
            for (int dx = -RAD; dx <= RAD; ++dx) hor += k[dx + RAD] * px_or_zero(in, W, H, x + dx, y + dy);
            num += k[dy + RAD] * hor;
        }
        if constexpr (RAD == 1) {  // intra.cl:3281-3285, 3437-3458
            const int nEdges = (x == 0) + (x == W - 1) + (y == 0) + (y == H - 1);
            const int k0 = k[0], k1 = k[1];
            den = nEdges >= 2 ? (k0 + 2 * k1 + k1 * k1) : (nEdges == 1 ? (2 * k0 + 3 * k1 + k1 * k1) : (4 * k0 + 4 * k1 + k1 * k1));
        } else {                   // intra.cl:3523-3551, 3753-3786
            auto ksum = [&](int i0, int j0) {
                int s = 0;
                for (int i = i0; i < 5; ++i)
                    for (int j = j0; j < 5; ++j) s += mip_k5(kidx, i - 2, j - 2);
                return s;
            };
            const bool oTB = (y == 0) || (y == H - 1), iTB = (y == 1) || (y == H - 2);
            const bool oLR = (x == 0) || (x == W - 1), iLR = (x == 1) || (x == W - 2);
            const bool oC = oTB && oLR, iC = iTB && iLR;
            const bool ifc = (oLR && iTB) || (iLR && oTB);
            const bool oE = !oC && !ifc && (oTB || oLR), iE = !iC && !ifc && (iTB || iLR);
            den = ksum(0, 0);
#pragma unroll
            for (int dy = -2; dy <= 2; ++dy)
                if (y + dy < 0 || y + dy >= H) den -= k[dy + 2];
            if (oC) den = ksum(2, 2);
            if (iC) den = ksum(1, 1);
            if (oE) den = ksum(0, 2);
            if (iE) den = ksum(0, 1);
            if (ifc) den = ksum(1, 2);
        }
    }
    out[(size_t)y * W + x] = (uint16_t)((num + den / 2) / den);
}

// ------------------------------------------------------------------------------------------
// Top-k (and, as k = 1, the stand-alone argmin): the k cheapest modes of every CU in ascending (cost, mode) order, from a cost table.
// One thread per CU; its 12/16/32 costs are one 16-byte aligned run (3/4/8 x LDG.128), kept in
// registers as unique keys (cost << 6 | mode) and selected k times.  Skipped CUs -> 0xFF / -1.
// ------------------------------------------------------------------------------------------
template <int MODES>
__device__ __forceinline__ void topk_cu(const int32_t* __restrict__ c, int k, uint8_t* __restrict__ om, int32_t* __restrict__ oc) {
    uint32_t key[MODES];
#pragma unroll
    for (int q = 0; q < MODES / 4; ++q) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(c) + q);
        key[4 * q] = ((uint32_t)v.x << 6) | (4 * q);
        key[4 * q + 1] = ((uint32_t)v.y << 6) | (4 * q + 1);
        key[4 * q + 2] = ((uint32_t)v.z << 6) | (4 * q + 2);
        key[4 * q + 3] = ((uint32_t)v.w << 6) | (4 * q + 3);
    }
    const bool skipped = (key[0] >> 6) == 0x03ffffffu;   // cost -1
    for (int j = 0; j < k; ++j) {
        uint32_t best = 0xffffffffu;
#pragma unroll
        for (int m = 0; m < MODES; ++m) best = min(best, key[m]);
#pragma unroll
        for (int m = 0; m < MODES; ++m) key[m] = key[m] == best ? 0xffffffffu : key[m];
        om[j] = skipped ? (uint8_t)0xFF : (uint8_t)(best & 63);
        oc[j] = skipped ? -1 : (int32_t)(best >> 6);
    }
}

__global__ void __launch_bounds__(256)
mip_topk_kernel(const int32_t* __restrict__ cost, int n_ctus, int k, uint8_t* __restrict__ modes_out, int32_t* __restrict__ costs_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_ctus * MIP_CUS_PER_CTU) return;
    const int ctu = idx / MIP_CUS_PER_CTU, cu = idx - ctu * MIP_CUS_PER_CTU;
    int lo = 0, hi = MIP_NUM_TYPES - 1;   // last type with cu_off <= cu
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)c_types[mid].cu_off <= cu) lo = mid; else hi = mid - 1;
    }
    const DevType& ty = c_types[lo];
    const int modes = ty.modes;
    const int32_t* c = cost + (size_t)ctu * MIP_COSTS_PER_CTU + ty.cost_off + (cu - ty.cu_off) * modes;
    uint8_t* om = modes_out + (size_t)idx * k;
    int32_t* oc = costs_out + (size_t)idx * k;
    if (modes == 12) topk_cu<12>(c, k, om, oc);
    else if (modes == 16) topk_cu<16>(c, k, om, oc);
    else topk_cu<32>(c, k, om, oc);
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
// the instantiations that exist: variant = OUT | COMPACT << 3 (the compact table excludes SAD / SATD)
constexpr int NUM_VARIANTS = 16;
static const void* cost_kernel_variant(int v) {
    switch (v) {
        case OUT_DEC: return (const void*)mip_cost_kernel<false, OUT_DEC>;
        case OUT_COST: return (const void*)mip_cost_kernel<false, OUT_COST>;
        case OUT_COST | OUT_DEC: return (const void*)mip_cost_kernel<false, OUT_COST | OUT_DEC>;
        case OUT_COST | OUT_SADSATD: return (const void*)mip_cost_kernel<false, OUT_COST | OUT_SADSATD>;
        case OUT_COST | OUT_SADSATD | OUT_DEC: return (const void*)mip_cost_kernel<false, OUT_COST | OUT_SADSATD | OUT_DEC>;
        case OUT_SADSATD: return (const void*)mip_cost_kernel<false, OUT_SADSATD>;
        case OUT_SADSATD | OUT_DEC: return (const void*)mip_cost_kernel<false, OUT_SADSATD | OUT_DEC>;
        case 8 | OUT_COST: return (const void*)mip_cost_kernel<true, OUT_COST>;
        case 8 | OUT_COST | OUT_DEC: return (const void*)mip_cost_kernel<true, OUT_COST | OUT_DEC>;
        default: return nullptr;
    }
}

cudaError_t kernels_init(const int* nchunks, const double* const* weights) {
    cudaError_t err;
    // the work list (mip_work_list.h: lane records, CU ordinals, the two chunk splits) and the CU type table
    static WorkList wl;
    if (!build_work_list(wl) || !split_work_list(wl, nchunks, weights)) return cudaErrorInvalidValue;
    DevType types[MIP_NUM_TYPES];
    memset(types, 0, sizeof(types));
    for (int t = 0; t < MIP_NUM_TYPES; ++t) {
        const mip_cu_type_t& s = MIP_TYPES[t];
        DevType& d = types[t];
        d.w = s.w; d.h = s.h; d.modes = s.modes;
        d.shape = wl.shape[t];
        d.cost_off = s.cost_off; d.cu_off = s.cu_off;
        d.parts_log2 = wl.parts_log2[t];
        d.narrow = (uint8_t)mip_compact_narrow(t);
        d.cmp_off = (uint32_t)mip_compact_type_offset(t);
    }
    static_assert(sizeof(LaneRec) == sizeof(uint2), "a lane record is a uint2 on the device");
    if (mip_compact_type_offset(MIP_NUM_TYPES) != MIP_COMPACT_BYTES_PER_CTU) return cudaErrorInvalidValue;
    if ((err = cudaMemcpyToSymbol(c_types, types, sizeof(types))) != cudaSuccess) return err;
    const std::vector<LaneRec> end_mark(32, LaneRec{0xffffffffu, 0u});
    for (int hf = 0; hf < 2; ++hf) {
        if ((err = cudaMemcpyToSymbol(g_lane, wl.lanes[hf].data(), wl.lanes[hf].size() * sizeof(LaneRec), (size_t)hf * MAX_WORK * 32 * sizeof(uint2))) != cudaSuccess) return err;
        if ((err = cudaMemcpyToSymbol(g_lane, end_mark.data(), 32 * sizeof(LaneRec), ((size_t)hf * MAX_WORK + MAX_WORK - 1) * 32 * sizeof(uint2))) != cudaSuccess) return err;
    }
    if ((err = cudaMemcpyToSymbol(c_chunk_begin, wl.begin, sizeof(wl.begin))) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(c_chunks, wl.chunks, sizeof(wl.chunks))) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(c_chunk_ord, wl.chunk_ord, sizeof(wl.chunk_ord))) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(c_ord2cu, wl.ord2cu, sizeof(wl.ord2cu))) != cudaSuccess) return err;
    const int* chunks_of = wl.chunks;
    // matrices: (coef - 32) as signed bytes in the padded shared-memory layout
    std::vector<uint8_t> mat(MAT_BYTES, 0);
    auto tpos = [](int p, int r) { return (p % r) * r + p / r; };   // output position of matrix row p in a transposed mode (intra.cl:485-487)
    for (int m = 0; m < 6; ++m)
        for (int p = 0; p < 64; ++p)
            for (int i = 0; i < 8; ++i) {
                const uint8_t v = (uint8_t)(int8_t)(mip_mat_id2(m, p, i) - 32);
                mat[M2_OFF + m * M2_STRIDE + p * 8 + i] = v;
                mat[M2T_OFF + m * M2_STRIDE + tpos(p, 8) * 8 + i] = v;
            }
    for (int m = 0; m < 8; ++m)
        for (int p = 0; p < 16; ++p)
            for (int i = 0; i < 8; ++i) {
                const uint8_t v = (uint8_t)(int8_t)(mip_mat_id1(m, p, i) - 32);
                mat[M1_OFF + m * M1_STRIDE + p * 8 + i] = v;
                mat[M1T_OFF + m * M1_STRIDE + tpos(p, 4) * 8 + i] = v;
            }
    for (int m = 0; m < 16; ++m)
        for (int p = 0; p < 16; ++p)
            for (int i = 0; i < 4; ++i) {
                const uint8_t v = (uint8_t)(int8_t)(mip_mat_id0(m, p, i) - 32);
                mat[M0_OFF + m * M0_STRIDE + p * 4 + i] = v;
                mat[M0T_OFF + m * M0_STRIDE + tpos(p, 4) * 4 + i] = v;
            }
    {   // the packed clamp of run_task() keeps (sum >> 6) + first in a signed 16-bit half: true for every bit depth up to 12
        // iff maxv * (largest row sum of |coef - 32|) / 64 + maxv + 1 < 32768, and the 4x accumulator stays far inside int32
        int worst = 0;
        for (int m = 0; m < 6; ++m)
            for (int p = 0; p < 64; ++p) { int t = 0; for (int i = 1; i < 8; ++i) t += abs(mip_mat_id2(m, p, i) - 32); worst = std::max(worst, t); }
        for (int m = 0; m < 8; ++m)
            for (int p = 0; p < 16; ++p) { int t = 0; for (int i = 0; i < 8; ++i) t += abs(mip_mat_id1(m, p, i) - 32); worst = std::max(worst, t); }
        if ((4095 * worst + 32) / 64 + 4096 >= 32768) return cudaErrorInvalidValue;
    }
    if ((err = cudaMemcpyToSymbol(g_mat, mat.data(), MAT_BYTES)) != cudaSuccess) return err;
    int ctas = 0;
    for (int v = 0; v < NUM_VARIANTS; ++v) {
        const void* fn = cost_kernel_variant(v);
        if (!fn) continue;
        if ((err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL)) != cudaSuccess) return err;
        if ((err = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return err;
        if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, fn, NT, SM_TOTAL)) != cudaSuccess) return err;
    }
    if (getenv("MIPB200_VERBOSE")) fprintf(stderr, "mipb200: cost kernel %d threads, %d B smem, %d CTA(s)/SM, %d / %d chunks per CTU half (throughput / lone frame)\n", NT, SM_TOTAL, ctas, chunks_of[0], chunks_of[1]);
    g_chunks[0] = chunks_of[0];
    g_chunks[1] = chunks_of[1];
    return cudaSuccess;
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

// A tensor map depends on nothing but (address, W, H), so encoded maps are kept: the engine's slots and a caller's
// resident frame pool hit the cache on every launch after their first (no driver call per frame).
struct MapCacheEntry { const uint16_t* ptr; int W, H; CUtensorMap map; };
static std::mutex g_map_mutex;
static MapCacheEntry g_map_cache[64];
static int g_map_count = 0, g_map_next = 0;

static cudaError_t make_frame_map(const uint16_t* d_frame, int W, int H, CUtensorMap* map) {
    {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        for (int i = 0; i < g_map_count; ++i)
            if (g_map_cache[i].ptr == d_frame && g_map_cache[i].W == W && g_map_cache[i].H == H) { *map = g_map_cache[i].map; return cudaSuccess; }
    }
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (!fn || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        g_encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t gstride[1] = {(cuuint64_t)W * sizeof(uint16_t)};
    const cuuint32_t box[2] = {STG_W, STG_H};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<uint16_t*>(d_frame), gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lk(g_map_mutex);
    MapCacheEntry& en = g_map_cache[g_map_next];
    en.ptr = d_frame; en.W = W; en.H = H; en.map = *map;
    g_map_next = (g_map_next + 1) % 64;
    if (g_map_count < 64) ++g_map_count;
    return cudaSuccess;
}

cudaError_t make_filter_params(int ft, int kidx, int bit_depth, FilterParams* fp) {
    memset(fp, 0, sizeof(*fp));
    fp->type = ft;
    fp->kidx = kidx;
    fp->full_den = 1;
    fp->full_magic = 0;
    if (ft == 0) return cudaSuccess;
    const bool is5 = ft >= 5, is2d = (ft == 3 || ft == 4 || ft == 7 || ft == 8);
    const int R = is5 ? 2 : 1, N = 2 * R + 1;
    int sum2d = 0;
    for (int dy = -R; dy <= R; ++dy)
        for (int dx = -R; dx <= R; ++dx) {
            const int k2 = is5 ? mip_k5(kidx, dy, dx) : mip_k3(kidx, dy, dx);
            const int k1 = (is5 ? mip_k5(kidx, -2, dy) : mip_k3(kidx, -1, dy)) * (is5 ? mip_k5(kidx, -2, dx) : mip_k3(kidx, -1, dx));
            fp->coef[(dy + R) * N + dx + R] = is2d ? k2 : k1;     // 1-D types: first row of the table, applied in x and in y
            sum2d += k2;
        }
    if (is2d || is5) fp->full_den = sum2d;                         // 1-D 5x5 takes its interior denominator from the 2-D table
    else { const int k0 = mip_k3(kidx, -1, -1), k1 = mip_k3(kidx, -1, 0); fp->full_den = 4 * k0 + 4 * k1 + k1 * k1; }
    fp->full_magic = (uint32_t)((1ull << 32) / fp->full_den + 1);
    // exactness of the multiply-high division over every numerator that can occur (weights x max sample + den/2)
    int wsum = 0;
    for (int i = 0; i < N * N; ++i) wsum += fp->coef[i];
    const uint64_t nmax = (uint64_t)wsum * ((1u << bit_depth) - 1) + fp->full_den / 2;
    for (uint64_t n = 0; n <= nmax; ++n)
        if (((n * fp->full_magic) >> 32) != n / fp->full_den) return cudaErrorInvalidValue;
    return cudaSuccess;
}

cudaError_t launch_costs(const uint16_t* d_frame, int W, int H, int bit_depth, const FilterParams& fp, int32_t* d_cost, int32_t* d_sad,
                         int32_t* d_satd, uint8_t* d_best_mode, int32_t* d_best_cost, bool lone_frame, bool compact, cudaStream_t st) {
    if (compact && (bit_depth > 10 || d_sad)) return cudaErrorInvalidValue;      // 16-bit entries hold 2 * 32 * 1023, not 2 * 32 * 4095
    if (bit_depth != 8 && bit_depth != 10 && bit_depth != 12) return cudaErrorInvalidValue;
    if ((d_best_mode == nullptr) != (d_best_cost == nullptr) || (d_sad == nullptr) != (d_satd == nullptr)) return cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(d_frame) & 15) != 0) return cudaErrorMisalignedAddress;   // TMA needs a 16-byte aligned frame
    CUtensorMap map;
    cudaError_t e = make_frame_map(d_frame, W, H, &map);
    if (e != cudaSuccess) return e;
    const int nctu = ((W + 127) >> 7) * ((H + 127) >> 7);
    const int split = lone_frame ? 1 : 0;
    const int grid = nctu * 2 * g_chunks[split];
    int maxv = (1 << bit_depth) - 1;
    const int variant = (d_cost ? OUT_COST : 0) | (d_sad ? OUT_SADSATD : 0) | (d_best_mode ? OUT_DEC : 0) | (compact ? 8 : 0);
    const void* fn = cost_kernel_variant(variant);
    if (!fn) return cudaErrorInvalidValue;           // no output at all, or SAD / SATD beside the compact table
    void* args[] = {&map, const_cast<FilterParams*>(&fp), &W, &H, const_cast<int*>(&split), &maxv, &d_cost, &d_sad, &d_satd, &d_best_mode, &d_best_cost};
    if ((e = cudaLaunchKernel(fn, dim3(grid), dim3(NT), args, SM_TOTAL, st)) != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t launch_filter(const uint16_t* d_in, uint16_t* d_out, int W, int H, int ft, int kidx, cudaStream_t st) {
    dim3 grid((W + 31) / 32, (H + 7) / 8), block(256);
    const bool is5 = ft >= 5, is2d = (ft == 3 || ft == 4 || ft == 7 || ft == 8);
    if (is2d && !is5) mip_filter_kernel<1, true><<<grid, block, 0, st>>>(d_in, d_out, W, H, kidx);
    else if (is2d) mip_filter_kernel<2, true><<<grid, block, 0, st>>>(d_in, d_out, W, H, kidx);
    else if (!is5) mip_filter_kernel<1, false><<<grid, block, 0, st>>>(d_in, d_out, W, H, kidx);
    else mip_filter_kernel<2, false><<<grid, block, 0, st>>>(d_in, d_out, W, H, kidx);
    return cudaGetLastError();
}

// Per-CU argmin = the shortlist of length 1 (same tie rule, same skipped-CU values, same [nCTU][5380] layout); with its
// 128-bit loads that kernel runs at 52 % of the HBM peak where a scalar-load argmin managed 20 %.
cudaError_t launch_decide(const int32_t* d_cost, int n_ctus, uint8_t* d_best_mode, int32_t* d_best_cost, cudaStream_t st) {
    return launch_topk(d_cost, n_ctus, 1, d_best_mode, d_best_cost, st);
}

cudaError_t launch_topk(const int32_t* d_cost, int n_ctus, int k, uint8_t* d_modes, int32_t* d_costs, cudaStream_t st) {
    if (k < 1 || k > MIP_TOPK_MAX) return cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(d_cost) & 15) != 0) return cudaErrorMisalignedAddress;   // 128-bit loads of a CU's run
    const int n = n_ctus * MIP_CUS_PER_CTU;
    mip_topk_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_cost, n_ctus, k, d_modes, d_costs);
    return cudaGetLastError();
}

}  // namespace mipb200
