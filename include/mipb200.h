/*
 * mipb200.h -- C ABI of the B200-native VVC MIP mode-decision engine.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference
 * (iagostorch/VVC-MIP-GPU) has no FFI layer: its host (main.cpp) drives four OpenCL kernels
 * through clSetKernelArg/clEnqueueNDRangeKernel.  Each entry point below names the
 * reference interface it replaces (file:line into the reference repository).
 *
 * Conventions: extern "C", plain pointers and sizes, no exceptions across the boundary.
 * Functions that touch the GPU switch to the engine's device and restore the calling thread's current device
 * before they return.
 * Every function returning int returns 0 on success and a negative MIPB200_E* code on
 * failure; mipb200_last_error() then holds a human-readable message (thread-local).
 * One engine drives one GPU.  An engine is not re-entrant; distinct engines may be driven
 * from distinct threads (one per GPU -- frames are independent, there is no collective).
 * There is NO CPU fallback: without a CUDA device mipb200_create() fails.
 */
#ifndef MIPB200_H
#define MIPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIPB200_COSTS_PER_CTU 97840 /* (CU,mode) pairs per 128x128 CTU, constants.h:1627-1629 */
#define MIPB200_CUS_PER_CTU 5380    /* CUs per CTU over the 47 CU types, constants.h:568-570 */
#define MIPB200_TOPK_MAX 12         /* the fewest modes any CU has (sizeId 2: 2 x 6 matrices) */
#define MIPB200_SKIPPED (-1)        /* cost of a CU that is not fully inside the frame
                                       (the reference leaves garbage there: intra.cl:96,232,717) */

/* error codes */
#define MIPB200_OK 0
#define MIPB200_EINVAL (-1)   /* bad argument / unsupported geometry */
#define MIPB200_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define MIPB200_ENODEV (-3)   /* no usable CUDA device */
#define MIPB200_EBUSY (-4)    /* submit with all slots in flight */
#define MIPB200_EEMPTY (-5)   /* collect with nothing in flight */
#define MIPB200_ENOMEM (-6)

/* what a frame's result carries (mipb200_config.emit bitmask) */
#define MIPB200_EMIT_COSTS 1u      /* int32 cost[nCTU][97840] = min(2*SAD, SATD), reference order */
#define MIPB200_EMIT_SAD_SATD 2u   /* int32 sad[...] and satd[...] in the same layout */
#define MIPB200_EMIT_DECISIONS 4u  /* uint8 best_mode[nCTU][5380] + int32 best_cost[nCTU][5380] */
#define MIPB200_EMIT_COSTS_COMPACT 8u /* instead of MIPB200_EMIT_COSTS: the same table, same order, 71 % of the bytes -- the
                                        costs of CU types of at most 32 samples (4x4, 8x4, 4x8) as uint16 (<= 65 472 by
                                        construction with samples of up to 10 bits; 0xFFFF = skipped), all others int32:
                                        mipb200_compact_bytes_per_ctu() = 276 672 bytes per CTU, mipb200_expand_costs()
                                        turns it into the int32 table.  The full-table rate is bound by the host link
                                        (56 MB per 1080p frame), so this is 1.4x the frames/s.  Not with bit_depth 12,
                                        MIPB200_EMIT_SAD_SATD or top_k > 1. */

/* filter_type: 0 = original samples (USE_ALTERNATIVE_SAMPLES 0, main.cpp:10);
 * 1..8 = the reference's availableFilters in order (constants.h:25-34):
 *   1 filterFrame_1d_int            2 filterFrame_1d_float
 *   3 filterFrame_2d_int_quarterCtu 4 filterFrame_2d_float_quarterCtu
 *   5 filterFrame_1d_int_5x5        6 filterFrame_1d_float_5x5
 *   7 filterFrame_2d_int_5x5_quarterCtu  8 filterFrame_2d_float_5x5_quarterCtu */
typedef struct mipb200_config {
    int width, height; /* luma size; width % 8 == 0, height % 4 == 0 (a superset of main.cpp:289-309; CUs that
                          cross the right or bottom frame edge are skipped: cost -1) */
    int device;        /* CUDA ordinal == --DeviceIndex (main.cpp:221-228) */
    int filter_type;   /* 0..8, see above (--FilterType, main.cpp:57) */
    int kernel_idx;    /* --KernelIdx (main.cpp:58): 0..4 for 3x3 filters, 0..2 for 5x5 */
    int slots;         /* frames in flight (>= 1; the reference has BUFFER_SLOTS 2, intra.cl:12) */
    unsigned emit;     /* MIPB200_EMIT_* bitmask, must not be 0 */
    int top_k;         /* 0 or 1: best mode only; 2..MIPB200_TOPK_MAX: with MIPB200_EMIT_DECISIONS the result also
                          carries the k cheapest modes of every CU (a shortlist for the encoder's RD search) */
    int bit_depth;     /* 0 or 10: the reference's hard-wired 10-bit pipeline (default sample 512 and clamp 1023:
                          intra.cl:61, 446, 482); 8 or 12 scale those constants to 1 << (bits - 1) and (1 << bits) - 1.
                          Samples must be < 1 << bits.  8-bit content run "as is" through the 10-bit pipeline, which is
                          what the reference does with it, is bit_depth 10. */
} mipb200_config;

typedef struct mipb200_engine mipb200_engine;

/* Result of one frame; pointers address the engine's pinned host ring and stay valid until
 * the next mipb200_collect()/mipb200_destroy() on the same engine -- mipb200_submit() calls in
 * between do not touch them (the ring holds one slot more than `slots`).  Unrequested outputs are
 * NULL.  Replaces return_minSadHad/SAD/SATD (main.cpp:660, main_aux_functions.h:585-630). */
typedef struct mipb200_result {
    int64_t poc;              /* tag given at submit */
    int n_ctus;               /* ceil(W/128) * ceil(H/128), raster order (intra.cl:44-45) */
    const int32_t* cost;      /* [n_ctus][97840] */
    const int32_t* sad;       /* [n_ctus][97840] */
    const int32_t* satd;      /* [n_ctus][97840] */
    const uint8_t* best_mode; /* [n_ctus][5380]; argmin over modes, lowest wins ties; 0xFF if skipped */
    const int32_t* best_cost; /* [n_ctus][5380] */
    float gpu_ms;             /* device time of this frame's kernels (CUDA events) */
    int top_k;                /* entries per CU in the two arrays below (0 when not requested) */
    const uint8_t* topk_mode; /* [n_ctus][5380][top_k]; ascending (cost, mode); 0xFF if skipped */
    const int32_t* topk_cost; /* [n_ctus][5380][top_k] */
    const void* cost_compact; /* [n_ctus][mipb200_compact_bytes_per_ctu()] with MIPB200_EMIT_COSTS_COMPACT (then cost is NULL) */
} mipb200_result;

/* Replaces the OpenCL platform/context/queue/buffer/program setup, main.cpp:87-315, 408-549. */
int mipb200_create(mipb200_engine** out, const mipb200_config* cfg);
void mipb200_destroy(mipb200_engine* e);

/* Pinned staging buffer (width*height uint16) of the slot the next submit will use.  Filling
 * it directly and passing the same pointer to mipb200_submit() avoids one host copy.
 * Returns NULL when every slot is in flight. */
uint16_t* mipb200_next_input(mipb200_engine* e);

/* Enqueue one frame: async H2D from the pinned slot, filter (if any), MIP costs, decisions,
 * async D2H.  `frame` = height*width uint16 row-major, samples 0..1023 (main.cpp:364-384);
 * Pageable memory is first copied into the slot's pinned buffer; a frame that already lives in
 * page-locked memory (mipb200_next_input(), mipb200_pin_host(), cudaHostAlloc/cudaHostRegister) is DMA'd in place and
 * must then stay alive and untouched until the frame has been collected.  Device or managed pointers are refused
 * (MIPB200_EINVAL; see mipb200_run_device).  Samples must be < 1 << bit_depth (not checked: larger values silently
 * overflow the packed 16-bit arithmetic -- the CLI checks its inputs).  Returns immediately.
 * Replaces one iteration of the frame loop, main.cpp:678-1241. */
int mipb200_submit(mipb200_engine* e, const uint16_t* frame, int64_t poc);

/* Wait for the oldest frame in flight and expose its results (FIFO).
 * Replaces clFinish + readMemobjsIntoArray_Distortion, main.cpp:1223-1250. */
int mipb200_collect(mipb200_engine* e, mipb200_result* out);

int mipb200_in_flight(const mipb200_engine* e);

/* How the frame's work is cut into thread blocks.  A frame that has the GPU to itself wants its last blocks short (the
 * frame's tail is then short: lowest latency of one frame); frames that overlap on the GPU -- several in flight, or
 * device-resident launches on several streams -- fill each other's tails and want few, equal blocks (fewest tile
 * stagings: highest frames/s).  AUTO (the default): mipb200_submit() uses the latency split when nothing else is in
 * flight and the throughput split otherwise; mipb200_run_device() uses the throughput split.  Results are identical
 * either way.  No reference counterpart (its NDRanges are fixed, main.cpp:1011-1200). */
#define MIPB200_LAUNCH_AUTO 0
#define MIPB200_LAUNCH_THROUGHPUT 1
#define MIPB200_LAUNCH_LATENCY 2
int mipb200_set_launch_mode(mipb200_engine* e, int mode);
int mipb200_num_ctus(int width, int height);

/* Number of CUDA devices (>= 0), or a negative MIPB200_E* code.  Replaces the reference's platform / device scan
 * (main.cpp:117-228: "COMPUTING ON GPU i" / "Incorrect GPU index. Only N GPUs are detected"). */
int mipb200_device_count(void);

/* Device-resident path: the frame is already in HBM and the results stay there (no host
 * copies).  All pointers are device pointers on the engine's GPU; any output may be NULL.
 * `stream` is a cudaStream_t (NULL = the engine's compute stream).  Asynchronous.  d_cost may be NULL when only the
 * decisions are wanted; d_best_mode and d_best_cost go together, and so do d_sad and d_satd; d_frame must be 16-byte aligned (TMA source).
 * ONE kernel: this is the fused equivalent of the reference's kernel sequence filterFrame_* ->
 * initBoundaries -> MIP_ReducedPred -> upsampleDistortion x3 (main.cpp:723-742, 819-844,
 * 925-946, 1011-1045, 1090-1124, 1167-1200). */
int mipb200_run_device(mipb200_engine* e, const uint16_t* d_frame, int32_t* d_cost, int32_t* d_sad,
                       int32_t* d_satd, uint8_t* d_best_mode, int32_t* d_best_cost, void* stream);

/* Only the low-pass filter: d_out[h][w] = filtered d_frame.  Same arguments as the
 * reference's filterFrame_* kernels (intra.cl:1639, 1828, 2311, 2539, 2856, 3042, 3267, 3508:
 * referenceFrame, filteredFrame, frameWidth, frameHeight, kernelIdx). */
int mipb200_filter_device(mipb200_engine* e, const uint16_t* d_frame, uint16_t* d_out, void* stream);

/* Per CU argmin over an existing cost table (device pointers; d_cost 16-byte aligned). */
int mipb200_decide_device(mipb200_engine* e, const int32_t* d_cost, uint8_t* d_best_mode,
                          int32_t* d_best_cost, void* stream);

/* The k (1..MIPB200_TOPK_MAX) cheapest modes of every CU of an existing cost table (device pointers, d_cost 16-byte
 * aligned): d_modes [nCTU][5380][k] uint8 and d_costs [nCTU][5380][k] int32 in ascending (cost, mode) order, i.e.
 * entry 0 equals mipb200_decide_device()'s answer.  No reference counterpart: the reference stops at the cost log
 * (main_aux_functions.h:735-798) and leaves the ranking to the consumer. */
int mipb200_topk_device(mipb200_engine* e, const int32_t* d_cost, int k, uint8_t* d_modes, int32_t* d_costs, void* stream);

/* Compact cost table -> int32 table (cost[n_ctus][97840], -1 for skipped CUs), on the host, with up to `threads` threads.
 * Layout of a CTU's compact record: the 47 type blocks in table order; a block holds modes * CUs entries in the order of the
 * int32 table; entries are uint16 when the type's CUs have at most 32 samples, int32 otherwise (csrc/mip_compact.h). */
size_t mipb200_compact_bytes_per_ctu(void);
int mipb200_expand_costs(const void* compact, int n_ctus, int32_t* cost, int threads);

/* Number of this library's kernels launched so far on this engine (bench.py: gpu_launches). */
long long mipb200_kernel_launches(const mipb200_engine* e);

/* Page-lock / release a host range the caller owns (cudaHostRegister), so that mipb200_submit() DMAs frames from
 * it in place instead of staging them through the slot's pinned buffer.  The reference uploads from pageable
 * memory (clEnqueueWriteBuffer from a malloc'd array, main.cpp:580-588, 886-898).  Needs a CUDA device. */
int mipb200_pin_host(void* ptr, size_t bytes);                 /* on the calling thread's current CUDA device */
int mipb200_pin_host_on(int device, void* ptr, size_t bytes);   /* on `device` (-1 = current); the caller's current device is untouched */
int mipb200_unpin_host(void* ptr);

/* Energy counter of the board behind CUDA ordinal `device`, in millijoules since the driver was loaded (NVML,
 * loaded at run time; MIPB200_ENODEV if NVML or the counter is unavailable).  Read it before and after a run to get
 * joules per frame -- what the reference obtains by integrating an nvidia-smi power trace over the stage stamps
 * (powerTracer_NVIDIA.py:9, computeEnergy_NVIDIA.py:44-96). */
int mipb200_device_energy_mj(int device, unsigned long long* millijoules);

/* Block until everything enqueued on the engine's streams has finished. */
int mipb200_sync(mipb200_engine* e);

const char* mipb200_last_error(void);
const char* mipb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MIPB200_H */
