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

// The same arithmetic in the same order (bit-identical results) for S <= 256 samples per ray, G = ceil(S / 32): every
// load of the ray is issued before the first use (G 512-byte raw rows + G 128-byte z rows in flight per warp instead of
// one), the depth of the next sample comes from the neighbouring lane instead of a second load, streaming cache hints on
// data that is read / written exactly once.  (The one-group-at-a-time kernel above measured 24 % of the HBM peak at
// 32,768 x 192 - profiles/r1_summary.md section 5: a dependent warp scan sat between consecutive loads.)
// rgb = sigmoid(raw[..., :3]) (:259) inside the weighted sum: ex2.approx + rcp.approx (each <= 2 ulp; the three channels of
// 192 samples were 45 % of the kernel's instructions with IEEE expf and division, and the kernel is issue-bound).  The
// density path (alpha = 1 - exp(-sigma dist), :246) keeps the accurate expf: 1 - exp(-x) amplifies its error for small x.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

template <int G>
__global__ void __launch_bounds__(256) r2l_raw2outputs_unrolled_kernel(const float* __restrict__ raw, const float* __restrict__ z_vals,
                                                                       const float* __restrict__ rays_d, int64_t n_rays, int n_samples,
                                                                       int white_bkgd, float* __restrict__ rgb_map,
                                                                       float* __restrict__ disp_map, float* __restrict__ acc_map,
                                                                       float* __restrict__ weights, float* __restrict__ depth_map) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const float4* rw = reinterpret_cast<const float4*>(raw + ray * n_samples * 4);
  const float* zv = z_vals + ray * n_samples;
  float4 r4[G];
  float z[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int i = 32 * g + lane;
    const bool in = i < n_samples;
    r4[g] = in ? __ldcs(rw + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    z[g] = in ? __ldcs(zv + i) : 0.f;
  }
  const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);        // torch.norm(rays_d) :255
  float carry = 1.f;
  float sr = 0.f, sg = 0.f, sb = 0.f, sdepth = 0.f, sacc = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int i = 32 * g + lane;
    const bool in = i < n_samples;
    float zn = __shfl_down_sync(0xffffffffu, z[g], 1);
    const float z_first_of_next = __shfl_sync(0xffffffffu, z[g + 1 < G ? g + 1 : g], 0);
    if (lane == 31) zn = z_first_of_next;
    float alpha = 0.f;
    if (in) {
      const float dist = ((i + 1 < n_samples) ? (zn - z[g]) : 1e10f) * dnorm;   // :249-257
      alpha = 1.f - expf(-fmaxf(r4[g].w, 0.f) * dist);                           // raw2alpha :246
    }
    const float t = in ? (1.f - alpha + 1e-10f) : 1.f;                           // :281-283
    const float incl = warp_incl_prod(t, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    const float w = alpha * (carry * excl);
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    if (in) {
      __stcs(weights + ray * n_samples + i, w);
      sr += w * fast_sigmoid(r4[g].x);                                           // sigmoid :259
      sg += w * fast_sigmoid(r4[g].y);
      sb += w * fast_sigmoid(r4[g].z);
      sdepth += w * z[g];
      sacc += w;
    }
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sdepth = warp_sum(sdepth); sacc = warp_sum(sacc);
  if (lane == 0) {
    const float q = sdepth / sacc;
    disp_map[ray] = (q != q) ? q : 1.f / fmaxf(1e-10f, q);     // NaN for a fully transparent ray, like the reference (see above)
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

// ----------------------------------------------------------------------------------------------
// Hierarchical resampling between the teacher's coarse and fine passes: inverse-CDF sampling of the coarse
// weights followed by the sorted merge with the coarse depths, one warp per ray, no host round trip.
// Reference: sample_pdf utils/run_nerf_raybased_helpers.py:283-330, call site utils/create_data.py:503-515 (which
// moves the tensors to the CPU and back: SURVEY.md row N1).
//   z_vals[N,S], weights[N,S] (coarse raw2outputs)  ->  z_samples[N,M], z_merged[N,S+M] (ascending)
//   u: uniforms, element (ray, j) at u[ray*u_stride + j]  (u_stride = 0: one shared row, e.g. linspace for det=True)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) r2l_sample_pdf_merge_kernel(const float* __restrict__ z_vals, const float* __restrict__ weights,
                                                                   const float* __restrict__ u, int64_t u_stride, int64_t n_rays,
                                                                   int S, int M, float* __restrict__ z_samples,
                                                                   float* __restrict__ z_merged, const float* __restrict__ bins_in) {
  extern __shared__ float smem_f[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int per_warp = 2 * S + S + M;
  float* cdf = smem_f + wib * per_warp;     // [S-1]
  float* bins = cdf + S;                    // [S-1]  z_vals_mid
  float* vals = bins + S;                   // [S+M]
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (ray >= n_rays) return;
  // two calling forms: (z_vals[N,S], weights[N,S]) as render_rays has them, or explicit (bins[N,S-1], weights[N,S-2])
  const float* z = bins_in ? nullptr : z_vals + ray * S;
  const float* w = bins_in ? weights + ray * (S - 2) : weights + ray * S + 1;   // weights[..., 1:-1]
  const int nb = S - 2;                     // number of pdf bins; cdf has nb + 1 = S - 1 entries, like z_vals_mid
  float part = 0.f;
  for (int b = lane; b < nb; b += 32) part += w[b] + 1e-5f;
  const float total = warp_sum(part);
  float carry = 0.f;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < nb; base += 32) {
    const int b = base + lane;
    float v = b < nb ? (w[b] + 1e-5f) / total : 0.f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += o;
    }
    if (b < nb) cdf[b + 1] = carry + v;
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
  for (int b = lane; b < S - 1; b += 32) bins[b] = bins_in ? bins_in[ray * (S - 1) + b] : .5f * (z[b + 1] + z[b]);
  if (z_merged)
    for (int i = lane; i < S; i += 32) vals[i] = z[i];
  __syncwarp();
  const int last = S - 2;                   // cdf.shape[-1] - 1
  for (int j = lane; j < M; j += 32) {
    const float uj = u[ray * u_stride + j];
    int lo = 0, hi = S - 1;                 // searchsorted(cdf, u, right=True): first index with cdf[idx] > u
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] > uj) hi = mid; else lo = mid + 1;
    }
    const int below = max(0, lo - 1), above = min(last, lo);
    const float c0 = cdf[below], c1 = cdf[above];
    float denom = c1 - c0;
    if (denom < 1e-5f) denom = 1.f;
    const float t = (uj - c0) / denom;
    const float smp = bins[below] + t * (bins[above] - bins[below]);
    z_samples[ray * M + j] = smp;
    vals[S + j] = smp;
  }
  __syncwarp();
  if (!z_merged) return;
  // Sorted merge by rank: element e of cat(z_vals, z_samples) goes to position #{k : (vals[k], k) < (vals[e], e)} (a total
  // order, so the ranks are a permutation and the result is torch.sort(cat(...)).values).  render_rays passes ascending
  // depths and - with perturb = 0 - ascending samples: then a rank is an index plus one binary search in the other list
  // (192 x ~8 probes per ray instead of 192 x 192 comparisons; this loop was 90 % of the kernel).  Whatever is not ascending
  // is counted by comparison as before.
  const int T = S + M;
  bool z_asc = true, s_asc = true;
  for (int i = lane; i + 1 < S; i += 32) z_asc = z_asc && (vals[i] <= vals[i + 1]);
  for (int j = lane; j + 1 < M; j += 32) s_asc = s_asc && (vals[S + j] <= vals[S + j + 1]);
  z_asc = __all_sync(0xffffffffu, z_asc);
  s_asc = __all_sync(0xffffffffu, s_asc);
  const float* zs = vals;          // [S]
  const float* sm = vals + S;      // [M]
  for (int e = lane; e < T; e += 32) {
    const float x = vals[e];
    int rank = 0;
    if (e < S) {
      // among the depths: the e before it when they ascend; among the samples (all have a larger index): those < x
      if (z_asc) rank = e;
      else for (int k = 0; k < S; ++k) { const float y = zs[k]; rank += (y < x) || (y == x && k < e); }
      if (s_asc) {
        int lo = 0, hi = M;        // lower bound: first j with sm[j] >= x
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm[mid] < x) lo = mid + 1; else hi = mid; }
        rank += lo;
      } else {
        for (int j = 0; j < M; ++j) rank += sm[j] < x;
      }
    } else {
      const int j0 = e - S;
      // among the depths (all have a smaller index): those <= x; among the samples: the j0 before it when they ascend
      if (z_asc) {
        int lo = 0, hi = S;        // upper bound: first k with zs[k] > x
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (zs[mid] <= x) lo = mid + 1; else hi = mid; }
        rank = lo;
      } else {
        for (int k = 0; k < S; ++k) rank += zs[k] <= x;
      }
      if (s_asc) rank += j0;
      else for (int j = 0; j < M; ++j) { const float y = sm[j]; rank += (y < x) || (y == x && j < j0); }
    }
    z_merged[ray * T + rank] = x;
  }
}

cudaError_t launch_sample_pdf_merge(const float* z_vals, const float* weights, const float* u, int64_t u_stride, int64_t n_rays,
                                    int S, int M, float* z_samples, float* z_merged, const float* bins_in, cudaStream_t stream) {
  const int warps = 8;
  const size_t smem = (size_t)warps * (3 * S + M) * sizeof(float);
  const int64_t blocks = (n_rays + warps - 1) / warps;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(r2l_sample_pdf_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  r2l_sample_pdf_merge_kernel<<<(unsigned)blocks, warps * 32, smem, stream>>>(z_vals, weights, u, u_stride, n_rays, S, M, z_samples,
                                                                              z_merged, bins_in);
  return cudaGetLastError();
}

cudaError_t launch_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int64_t n_rays, int n_samples,
                               int white_bkgd, float* rgb_map, float* disp_map, float* acc_map, float* weights,
                               float* depth_map, cudaStream_t stream) {
  const int warps = 8;
  const int64_t blocks = (n_rays + warps - 1) / warps;
#define R2L_R2O(G)                                                                                                              \
  case G:                                                                                                                      \
    r2l_raw2outputs_unrolled_kernel<G><<<(unsigned)blocks, warps * 32, 0, stream>>>(raw, z_vals, rays_d, n_rays, n_samples,     \
                                                                                    white_bkgd, rgb_map, disp_map, acc_map,    \
                                                                                    weights, depth_map);                       \
    break;
  switch ((n_samples + 31) / 32) {
    R2L_R2O(1) R2L_R2O(2) R2L_R2O(3) R2L_R2O(4) R2L_R2O(5) R2L_R2O(6) R2L_R2O(7) R2L_R2O(8)
    default:      // more than 256 samples per ray: one group at a time
      r2l_raw2outputs_kernel<<<(unsigned)blocks, warps * 32, 0, stream>>>(raw, z_vals, rays_d, n_rays, n_samples, white_bkgd, rgb_map,
                                                                        disp_map, acc_map, weights, depth_map);
  }
#undef R2L_R2O
  return cudaGetLastError();
}

cudaError_t launch_embed(const float* x, float* out, int64_t n, int dim, int L, int style, cudaStream_t stream) {
  const int64_t total = n * dim * (int64_t)(L + 1);
  r2l_embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, out, n, dim, L, style);
  return cudaGetLastError();
}

}  // namespace r2l
