// img2mse and its gradient in one pass (HBM-bound, 24 B read + 12..16 B written per ray).
// Reference: loss_rgb = img2mse(rgb[:, :3], target_s[:, :3]) * args.lw_rgb with img2mse = mean((x - y)^2)
// (model/nerf_raybased.py:18, main.py:1377), whose autograd gives dL/drgb = lw * 2 (rgb - t) / (3 N); and the per-ray error
// torch.mean((rgb - target_s)^2, dim=1) the hard-example pool sorts by (main.py:1411-1413).  The reference runs these as
// ~8 elementwise/reduction launches plus a psnr.item() host sync per step; here they are one launch and no sync.
#include "kernels.cuh"

namespace r2l {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 128;

// scratch: [kLossMaxBlocks] float partial sums, then one int ticket (must be zero before the first launch; the kernel
// leaves it zero).  The last block to finish adds the partials in block order: the loss is bit-reproducible.
__global__ void __launch_bounds__(kLossThreads) r2l_mse_loss_grad_kernel(const float* __restrict__ rgb, const float* __restrict__ target,
                                                                        int64_t n, int target_stride, float grad_scale, float loss_scale,
                                                                        float* __restrict__ grad_rgb, float* __restrict__ per_ray,
                                                                        float* __restrict__ loss, float* scratch) {
  __shared__ float warp_sums[kLossThreads / 32];
  __shared__ int is_last;
  float acc = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * kLossThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kLossThreads) {
    float e = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = rgb[3 * r + c] - target[(int64_t)target_stride * r + c];
      if (grad_rgb) grad_rgb[3 * r + c] = d * grad_scale;
      e = fmaf(d, d, e);
    }
    if (per_ray) per_ray[r] = e * (1.f / 3.f);
    acc += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) s += warp_sums[w];
    scratch[blockIdx.x] = s;
    __threadfence();
    int* ticket = reinterpret_cast<int*>(scratch + kLossMaxBlocks);
    is_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    if (is_last) {
      __threadfence();
      float total = 0.f;
      for (unsigned b = 0; b < gridDim.x; ++b) total += __ldcg(scratch + b);
      *loss = total * loss_scale;
      *ticket = 0;
    }
  }
}

cudaError_t launch_mse_loss_grad(const float* rgb, const float* target, int64_t n, int target_stride, float grad_scale, float loss_scale,
                                 float* grad_rgb, float* per_ray, float* loss, float* scratch, cudaStream_t stream) {
  int64_t blocks = (n + kLossThreads - 1) / kLossThreads;
  if (blocks > kLossMaxBlocks) blocks = kLossMaxBlocks;
  if (blocks < 1) blocks = 1;
  r2l_mse_loss_grad_kernel<<<(int)blocks, kLossThreads, 0, stream>>>(rgb, target, n, target_stride, grad_scale, loss_scale, grad_rgb, per_ray, loss, scratch);
  return cudaGetLastError();
}

}  // namespace r2l
