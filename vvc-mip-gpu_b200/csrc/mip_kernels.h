// Internal launcher API of the CUDA kernels (the C ABI lives in mip_engine.cu / include/mipb200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mipb200 {

// One-time per device: packs the MIP matrices, the CU tables and the work list into device memory.
cudaError_t kernels_init(int chunks_per_ctu);
int kernels_chunks_per_ctu();

// Fused MIP cost kernel for one frame.  d_orig: samples the distortion is measured against;
// d_ref: samples the boundaries come from (== d_orig for original-sample mode).
// cost/sad/satd: [nCTU][97840] int32, sad/satd may be null.
cudaError_t launch_costs(const uint16_t* d_orig, const uint16_t* d_ref, int W, int H, int32_t* d_cost,
                         int32_t* d_sad, int32_t* d_satd, cudaStream_t st);

// Low-pass filter of a whole frame (alternative samples), filter_type 1..8.
cudaError_t launch_filter(const uint16_t* d_in, uint16_t* d_out, int W, int H, int filter_type, int kernel_idx,
                          cudaStream_t st);

// Per-CU argmin over the cost table.
cudaError_t launch_decide(const int32_t* d_cost, int n_ctus, uint8_t* d_best_mode, int32_t* d_best_cost,
                          cudaStream_t st);

}  // namespace mipb200
