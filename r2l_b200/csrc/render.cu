// HBM-bound companions of the MLP kernels: volumetric compositing (raw2outputs) and dense positional encodings.
// Reference: /root/reference/model/nerf_raybased.py raw2outputs :226-295 (identical copy utils/create_data.py:335-402),
// PositionalEmbedder.__call__ :198-208, Embedder.embed :54-55.
#include "kernels.cuh"

namespace r2l {

// ----------------------------------------------------------------------------------------------
// raw2outputs: one warp per ray; lane i owns samples i, i+32, ...; transmittance = exclusive product scan.
// Algorithmic bytes per ray: 4*(4S + S + 3) in, 4*(6 + S) out.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v *= o;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__global__ void __launch_bounds__(256) r2l_raw2outputs_kernel(const float* __restrict__ raw, const float* __restrict__ z_vals,
                                                              const float* __restrict__ rays_d, int64_t n_rays, int n_samples,
                                                              int white_bkgd, float* __restrict__ rgb_map,
                                                              float* __restrict__ disp_map, float* __restrict__ acc_map,
                                                              float* __restrict__ weights, float* __restrict__ depth_map) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const float* rw = raw + ray * n_samples * 4;
  const float* zv = z_vals + ray * n_samples;
  const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);        // torch.norm(rays_d) :255
  float carry = 1.f;                                               // transmittance entering this 32-sample group
  float sr = 0.f, sg = 0.f, sb = 0.f, sdepth = 0.f, sacc = 0.f;
  for (int base = 0; base < n_samples; base += 32) {
    const int i = base + lane;
    const bool in = i < n_samples;
    float alpha = 0.f, z = 0.f;
    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) {
      r4 = *reinterpret_cast<const float4*>(rw + 4 * i);
      z = zv[i];
      const float zn = (i + 1 < n_samples) ? zv[i + 1] : 0.f;
      const float dist = ((i + 1 < n_samples) ? (zn - z) : 1e10f) * dnorm;   // :249-257
      alpha = 1.f - expf(-fmaxf(r4.w, 0.f) * dist);                            // raw2alpha :246
    }
    const float t = in ? (1.f - alpha + 1e-10f) : 1.f;                         // :281-283
    const float incl = warp_incl_prod(t, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    const float w = alpha * (carry * excl);
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    if (in) {
      weights[ray * n_samples + i] = w;
      sr += w / (1.f + expf(-r4.x));                                           // sigmoid :259
      sg += w / (1.f + expf(-r4.y));
      sb += w / (1.f + expf(-r4.z));
      sdepth += w * z;
      sacc += w;
    }
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sdepth = warp_sum(sdepth); sacc = warp_sum(sacc);
  if (lane == 0) {
    const float q = sdepth / sacc;
    // torch.max(1e-10, q) propagates NaN (0/0 for a fully transparent ray); reproduce, do not "fix" 
    disp_map[ray] = (q != q) ? q : 1.f / fmaxf(1e-10f, q);
    acc_map[ray] = sacc;
    depth_map[ray] = sdepth;
    const float bg = white_bkgd ? (1.f - sacc) : 0.f;
    rgb_map[ray * 3 + 0] = sr + bg;
    rgb_map[ray * 3 + 1] = sg + bg;
    rgb_map[ray * 3 + 2] = sb + bg;
  }
}

// ----------------------------------------------------------------------------------------------
// dense positional encodings (only for callers that really want the tensor; the fused MLP kernels never do)
//   style 0 (PositionalEmbedder, R2L): per coordinate [sin(x f_0..f_{L-1}), cos(...), x]   -> [N, D*(2L+1)]
//   style 1 (Embedder, teacher)      : [x (D), sin(x f_0) (D), cos(x f_0) (D), sin(x f_1) ...]  -> [N, D*(2L+1)]
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) r2l_embed_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n,
                                                        int dim, int L, int style) {
  const int E = 2 * L + 1;
  const int64_t total = n * dim * (int64_t)(L + 1);   // one thread per (row, coord, frequency | identity)
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int f = (int)(idx % (L + 1));
  const int c = (int)((idx / (L + 1)) % dim);
  const int64_t row = idx / ((int64_t)(L + 1) * dim);
  const float v = x[row * dim + c];
  float* o = out + row * (int64_t)dim * E;
  if (f == L) {
    if (style == 0) o[c * E + 2 * L] = v; else o[c] = v;
    return;
  }
  float s, co;
  sincosf(__fmul_rn(v, (float)(1 << f)), &s, &co);
  if (style == 0) {
    o[c * E + f] = s;
    o[c * E + L + f] = co;
  } else {
    o[dim + (2 * f) * dim + c] = s;
    o[dim + (2 * f + 1) * dim + c] = co;
  }
}

cudaError_t launch_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int64_t n_rays, int n_samples,
                               int white_bkgd, float* rgb_map, float* disp_map, float* acc_map, float* weights,
                               float* depth_map, cudaStream_t stream) {
  const int warps = 8;
  const int64_t blocks = (n_rays + warps - 1) / warps;
  r2l_raw2outputs_kernel<<<(unsigned)blocks, warps * 32, 0, stream>>>(raw, z_vals, rays_d, n_rays, n_samples, white_bkgd, rgb_map,
                                                                    disp_map, acc_map, weights, depth_map);
  return cudaGetLastError();
}

cudaError_t launch_embed(const float* x, float* out, int64_t n, int dim, int L, int style, cudaStream_t stream) {
  const int64_t total = n * dim * (int64_t)(L + 1);
  r2l_embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, out, n, dim, L, style);
  return cudaGetLastError();
}

}  // namespace r2l
