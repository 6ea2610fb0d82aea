// Host-visible launch interface of the CUDA kernels (internal to the library; the public ABI is include/r2l_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.cuh"

namespace r2l {

struct FwdParams {
  const float* in0;       // rays_o[N,3] | pts[N,48] | x[N,1008]
  const float* in1;       // rays_d[N,3] | unused
  const float* t_rand;    // [N,16] or nullptr (kInputRays only)
  float z_lo[kSamples];   // z_vals (no jitter) or `lower` (jitter)
  float z_diff[kSamples]; // `upper - lower` (jitter)
  const uint8_t* packed;
  float* rgb;             // [N,3]
  float* h_scratch;       // [gridDim.x][128][256] fp32: head output kept for the outer residual
  long long* stats;       // optional [gridDim.x][8] cycle counters (debug), nullptr in production
  int64_t n_rays;
  int num_tiles;
  int input_kind;
};

cudaError_t launch_pack(const float* params, void* packed, cudaStream_t stream);
cudaError_t launch_fwd(const FwdParams& p, int grid, cudaStream_t stream);
cudaError_t launch_umma_selftest(const float* A, const void* images, float* C, cudaStream_t stream);

}  // namespace r2l
