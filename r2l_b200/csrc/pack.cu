// Weight packing: flat fp32 state_dict buffer -> fp16 hi/lo UMMA operand images (of kWeightScale * W) + fp32 side tables.
// Runs once per parameter update (inference: once; training: once per optimizer step).
#include "kernels.cuh"
#include "ptx.cuh"

namespace r2l {

// fp32 side tables (biases, running sums of the second-layer biases, tail); one block, the first row of the
// pack_images grid so that its serial 43-step sum hides behind the image blocks
__device__ __forceinline__ void pack_tables(const float* __restrict__ params, uint8_t* __restrict__ packed) {
  const int col = threadIdx.x;
  float* cum = reinterpret_cast<float*>(packed + kPackOffCumBias);
  float* headb = reinterpret_cast<float*>(packed + kPackOffHeadB);
  float* b1 = reinterpret_cast<float*>(packed + kPackOffB1);
  float* tailw = reinterpret_cast<float*>(packed + kPackOffTailW);
  float* tailb = reinterpret_cast<float*>(packed + kPackOffTailB);
  float* b2t = reinterpret_cast<float*>(packed + kPackOffB2);
  // every load first (independent, read-only path): the block's run time is one memory round trip, not 43
  float b2[kBlocks], b1v[kBlocks];
#pragma unroll
  for (int k = 0; k < kBlocks; ++k) {
    b2[k] = __ldg(params + off_body_b(2 * k + 1) + col);
    b1v[k] = __ldg(params + off_body_b(2 * k) + col);
  }
  const float hb = __ldg(params + kOffHeadB + col);
  float tw[kOutDim];
#pragma unroll
  for (int c = 0; c < kOutDim; ++c) tw[c] = __ldg(params + kOffTailW + c * kWidth + col);
  float acc = 0.f;
  cum[col] = 0.f;
#pragma unroll
  for (int k = 0; k < kBlocks; ++k) {
    acc = __fadd_rn(acc, b2[k]);
    cum[(k + 1) * kWidth + col] = acc;
    b1[k * kWidth + col] = b1v[k];
    b2t[k * kWidth + col] = b2[k];
  }
  headb[col] = hb;
#pragma unroll
  for (int c = 0; c < kOutDim; ++c) tailw[c * kWidth + col] = tw[c];
  if (col < 4) tailb[col] = col < kOutDim ? __ldg(params + kOffTailB + col) : 0.f;
}

// One thread = one 16-byte swizzle unit (8 consecutive k) of one (n-row) of one image PAIR (hi+lo).
__global__ void __launch_bounds__(256) pack_images_kernel(const float* __restrict__ params,
                                                          uint8_t* __restrict__ packed) {
  if (blockIdx.y == 0) {                                     // first grid row (dispatched first): the tables
    if (blockIdx.x == 0) pack_tables(params, packed);
    return;
  }
  const int ip = blockIdx.y - 1;                             // image pair
  const int unit = blockIdx.x * blockDim.x + threadIdx.x;    // 0 .. 2047
  const int n = unit >> 3, j = unit & 7;
  float v[8];
  if (ip < 32) {
    const bool fused = ip < 16;
    const int chunk = fused ? ip : ip - 16;
    const float* w = params + kOffHeadW + (int64_t)n * kInDim;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int t = 8 * j + e;
      const int f = fused ? fused_slot_to_feature(chunk, t) : (64 * chunk + t < kInDim ? 64 * chunk + t : -1);
      v[e] = f >= 0 ? __ldg(w + f) : 0.f;
    }
  } else if (ip < 32 + 4 * kBodyLayers) {
    const int l = (ip - 32) >> 2, c = (ip - 32) & 3;
    const float* w = params + off_body_w(l) + (int64_t)n * kWidth + 64 * c + 8 * j;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __ldg(w + e);
  } else {
    const int idx = ip - 32 - 4 * kBodyLayers;
    const int q = idx >> 2, c = idx & 3;
    const int blk = (kBlocks - 1) - (q >> 1);
    const int l = 2 * blk + ((q & 1) ? 0 : 1);   // W2 of the block first, then W1
    const float* w = params + off_body_w(l) + n;  // B[n][kk] = W[kk][n]
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __ldg(w + (int64_t)(64 * c + 8 * j + e) * kWidth);
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split2(v[2 * e] * kWeightScale, v[2 * e + 1] * kWeightScale, hi[e], lo[e]);   // see ptx.cuh
  const uint32_t off = sw128_offset(n, 8 * j);
  uint8_t* img = packed + (int64_t)(2 * ip) * kWImageBytes;
  *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(img + kWImageBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

cudaError_t launch_pack(const float* params, void* packed, cudaStream_t stream) {
  dim3 grid(2048 / 256, kNumImages / 2 + 1);
  pack_images_kernel<<<grid, 256, 0, stream>>>(params, static_cast<uint8_t*>(packed));
  return cudaGetLastError();
}

}  // namespace r2l
