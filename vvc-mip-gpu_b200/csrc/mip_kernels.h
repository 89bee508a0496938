// Internal launcher API of the CUDA kernels (the C ABI lives in mip_engine.cu / include/mipb200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mipb200 {

// One-time per device: packs the MIP matrices, the CU tables and the work list into device memory.
// Two splits of a CTU half's work into chunks (= CTAs): [0] used when frames overlap on the GPU, [1] for a lone frame.
// nchunks[sp] = chunks per half, weights[sp] (may be null: equal) = relative cost share of each chunk, in launch order.
// cudaErrorInvalidValue: a chunk would hold more CUs than the shared-memory decision table (use more chunks).
cudaError_t kernels_init(const int* nchunks, const double* const* weights);

// Low-pass filter of the fused path, prepared once per engine and passed to the kernel by value (constant-bank
// operands): the (2R+1)^2 tap weights (k (x) k for the 1-D types), the denominator of a sample whose whole window
// lies inside the frame, and the multiplier that turns the division by it into a multiply-high.
struct FilterParams {
    int type;             // 0 none, 1..8 = availableFilters order
    int kidx;
    int coef[25];         // row-major (2R+1) x (2R+1)
    int full_den;
    uint32_t full_magic;  // floor(n / full_den) == (n * full_magic) >> 32 for every numerator that can occur (checked)
};
cudaError_t make_filter_params(int filter_type, int kernel_idx, int bit_depth, FilterParams* fp);

// Fused MIP kernel for one frame: TMA-staged tiles, optional low-pass filter of the reference samples
// (filter_type 0 = original samples, 1..8 = availableFilters order) applied in shared memory, boundaries,
// reduced prediction, up-sampling, SAD/SATD.  cost/sad/satd: [nCTU][97840] int32, sad/satd may be null.
// d_frame must be 16-byte aligned.
// d_best_mode/d_best_cost (both or neither): per-CU argmin over the modes, produced by the same kernel.
// d_cost may be null when only the decisions are wanted.
// bit_depth: 10 = the reference (clamp 1023, default sample 512: intra.cl:61, 446, 482); 8 and 12 scale those constants.
// lone_frame: nothing else will share the GPU with this launch -> the split with the short tail (see kernels_init).
cudaError_t launch_costs(const uint16_t* d_frame, int W, int H, int bit_depth, const FilterParams& fp, int32_t* d_cost,
                         int32_t* d_sad, int32_t* d_satd, uint8_t* d_best_mode, int32_t* d_best_cost, bool lone_frame, bool compact,
                         cudaStream_t st);   // compact: d_cost is a compact table (mip_compact.h), bit depth <= 10, no SAD / SATD

// Low-pass filter of a whole frame (alternative samples), filter_type 1..8.
cudaError_t launch_filter(const uint16_t* d_in, uint16_t* d_out, int W, int H, int filter_type, int kernel_idx,
                          cudaStream_t st);

// Per-CU argmin over the cost table.
cudaError_t launch_decide(const int32_t* d_cost, int n_ctus, uint8_t* d_best_mode, int32_t* d_best_cost,
                          cudaStream_t st);

// The k cheapest modes of every CU, ascending (cost, mode): modes [nCTU][5380][k] u8, costs [nCTU][5380][k] int32.
#define MIP_TOPK_MAX 12   // the fewest modes any CU has (sizeId 2: 2 x 6 matrices)
cudaError_t launch_topk(const int32_t* d_cost, int n_ctus, int k, uint8_t* d_modes, int32_t* d_costs, cudaStream_t st);

}  // namespace mipb200
