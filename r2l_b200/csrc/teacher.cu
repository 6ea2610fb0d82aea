// Teacher NeRF query: points + view directions -> raw (rgb, sigma), the 8x256 MLP with skip connection and view
// branch as ONE persistent tcgen05 kernel per 128-point tile.  Embeddings are built in registers, activations live
// in shared memory / TMEM, only 24 B/point go in and 16 B/point come out.
//
// Reference: /root/reference/model/nerf_raybased.py  Embedder.embed :54-55 (get_embedder :58-73),
// run_network :312-334, NeRF.__init__ :339-375, NeRF.forward :377-401 (use_viewdirs=True, D=8, W=256, skips=[4]);
// call sites utils/create_data.py:490,521.
//
// Layer program (weights streamed as 32 KiB fp16 hi/lo K-major images, N padded to 256):
//   T0  pts63 (1 chunk)                -> 256, relu          T1..T4  256 -> 256, relu
//   T5  [h256, pts63] (5 chunks)       -> 256, relu          T6, T7  256 -> 256, relu      (alpha = w_a . h7 + b_a on CUDA cores)
//   T8  feature: 256 -> 256 (linear)   T9  [feature256, dirs27] (5 chunks) -> 128, relu   (rgb = W_rgb hv + b on CUDA cores)
// Warp roles and barriers are those of chain.cu (producer / MMA issuer / TMEM allocator / 8 epilogue warps).
#include "kernels.cuh"
#include "ptx.cuh"

namespace r2l {

constexpr int kTNumWStages = 3;
constexpr int kTThreads = 384;
constexpr int kTEpiWarps = 8;
constexpr uint32_t kTSmemA = 0;
constexpr uint32_t kTSmemW = kABytes;
constexpr uint32_t kTSmemBar = kTSmemW + kTNumWStages * kWImageBytes;
// The 128 x 4 floats the two column halves exchange at the end of a tile live in A slot 1: every GEMM of the tile has
// completed by then, and slots 1-3 are next written by these same warps after the next tile's first layer.
constexpr uint32_t kTSmemOut = kTSmemA + kAChunkBytes;
constexpr uint32_t kTSmemBytes = kTSmemBar + 256 + 1024;
static_assert(kTSmemBytes <= 232448, "exceeds the 227 KiB dynamic shared memory limit");

enum : uint32_t {
  kTBarWFull = 0,
  kTBarWEmpty = kTBarWFull + kTNumWStages,
  kTBarAFull = kTBarWEmpty + kTNumWStages,   // [4]
  kTBarAEmpty = kTBarAFull + kAChunks,       // slot 0 only (5-chunk layers)
  kTBarAccFull = kTBarAEmpty + 1,
  kTBarCount
};

__constant__ int kTeacherChunks[kTeacherLayers] = {1, 4, 4, 4, 4, 5, 4, 4, 4, 5};

// feature index of fused slot `slot` in the reference's Embedder layout [x(3), sin f0 (3), cos f0 (3), ...]
__host__ __device__ inline int teacher_slot_to_feature(int slot, int nf) {
  if (slot < 6 * nf) {
    const int p = slot >> 1, c = p / nf, f = p % nf;
    return 3 + 6 * f + ((slot & 1) ? 3 : 0) + c;
  }
  if (slot < 6 * nf + 3) return slot - 6 * nf;
  return -1;
}

// fused-order encoding of a 3-vector with NF frequencies into 32 of the 64 slots of a K chunk (see layout.cuh:
// pair p = c*NF + f -> slots 2p (sin), 2p+1 (cos); then the three raw coordinates; zeros after)
template <int NF>
__device__ __forceinline__ void encode_generic(const float (&x)[3], uint32_t hf, float (&out)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) out[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int p = (int)hf * 16 + i;
    if (p < 3 * NF) {
      const int c = p / NF, f = p % NF;
      float s, co;
      sincosf(__fmul_rn(x[c], static_cast<float>(1 << f)), &s, &co);
      out[2 * i] = s;
      out[2 * i + 1] = co;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int slot = 6 * NF + c;
    if ((slot >> 5) == (int)hf) out[slot & 31] = x[c];
  }
}

__device__ __forceinline__ void t_store_a_half(uint32_t a_chunk_addr, uint32_t row, uint32_t hf, const float (&v)[32]) {
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2(v[8 * jj + 2 * e], v[8 * jj + 2 * e + 1], hi[e], lo[e]);
    const uint32_t off = row * 128u + (((4u * hf + jj) ^ (row & 7u)) << 4);
    st_shared_v4(a_chunk_addr + off, hi[0], hi[1], hi[2], hi[3]);
    st_shared_v4(a_chunk_addr + kPlaneBytes + off, lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(kTThreads, 1) r2l_teacher_kernel(const __grid_constant__ TeacherParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + kTSmemBar;
  auto bar = [&](uint32_t i) { return bar0 + 8u * i; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kTSmemBar + 8 * kTBarCount);
  float* out_smem = reinterpret_cast<float*>(smem_gen + kTSmemOut);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kTNumWStages; ++i) {
      mbar_init(bar(kTBarWFull + i), 1);
      mbar_init(bar(kTBarWEmpty + i), 1);
    }
    for (int i = 0; i < kAChunks; ++i) mbar_init(bar(kTBarAFull + i), kTEpiWarps);
    mbar_init(bar(kTBarAEmpty), 1);
    mbar_init(bar(kTBarAccFull), 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int i = 0; i < 2 * kTeacherImagePairs; ++i, ++it) {
          const uint32_t ws = it % kTNumWStages, ph = (it / kTNumWStages) & 1u;
          mbar_wait(bar(kTBarWEmpty + ws), ph ^ 1u);
          mbar_arrive_expect_tx(bar(kTBarWFull + ws), kWImageBytes);
          bulk_g2s(smem_base + kTSmemW + ws * kWImageBytes, p.packed + (int64_t)i * kWImageBytes, kWImageBytes, bar(kTBarWFull + ws));
        }
      }
    }
  } else if (warp == 1) {
    {   // the whole warp waits, one elected lane issues (keeps the descriptors in uniform registers, see chain.cu)
      constexpr uint32_t idesc = umma_idesc_f16(128, 256, 0, 0);
      uint32_t it = 0, a_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 0; l < kTeacherLayers; ++l) {
          const int nkc = kTeacherChunks[l];
          const uint32_t d = tmem_base + 256u * (l & 1);
          for (int kc = 0; kc < nkc; ++kc) {
            const uint32_t slot = kc & 3;
            mbar_wait(bar(kTBarAFull + slot), (a_phase >> slot) & 1u);
            a_phase ^= 1u << slot;
            const uint32_t a_hi = smem_base + kTSmemA + slot * kAChunkBytes, a_lo = a_hi + kPlaneBytes;
            {
              const uint32_t ws = it % kTNumWStages;
              mbar_wait(bar(kTBarWFull + ws), (it / kTNumWStages) & 1u);
              tc_fence_after_sync();
              const uint32_t b = smem_base + kTSmemW + ws * kWImageBytes;
              if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_f16(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b + 32 * ks, 16, 1024), idesc,
                            (kc == 0 && ks == 0) ? 0u : 1u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_f16(d, umma_desc_sw128(a_lo + 32 * ks, 16, 1024), umma_desc_sw128(b + 32 * ks, 16, 1024), idesc, 1u);
                umma_commit(bar(kTBarWEmpty + ws));
              }
              ++it;
            }
            {
              const uint32_t ws = it % kTNumWStages;
              mbar_wait(bar(kTBarWFull + ws), (it / kTNumWStages) & 1u);
              tc_fence_after_sync();
              const uint32_t b = smem_base + kTSmemW + ws * kWImageBytes;
              if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_f16(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b + 32 * ks, 16, 1024), idesc, 1u);
                umma_commit(bar(kTBarWEmpty + ws));
                if (nkc == 5 && kc == 0) umma_commit(bar(kTBarAEmpty));   // slot 0 is recycled for the 5th chunk
              }
              ++it;
            }
          }
          if (elect_one_sync()) umma_commit(bar(kTBarAccFull));
        }
      }
    }
  } else if (warp >= 4) {
    const uint32_t ew = warp - 4;
    const uint32_t q = ew & 3u, hf = ew >> 2;
    const uint32_t row = q * 32u + lane;
    const uint32_t tmem_row = tmem_base + ((q * 32u) << 16);
    const float* tab = reinterpret_cast<const float*>(p.packed + kTeacherPackOffTables);
    const float* biases = tab;                                   // [10][256]: T0..T7, feature, views(128, zero padded)
    const float* alpha_w = tab + kTeacherLayers * kWidth;        // [256]
    const float* rgb_w = alpha_w + kWidth;                       // [3][128]
    const float* scal = rgb_w + 3 * 128;                         // alpha_b, rgb_b[3]
    uint32_t acc_phase = 0;

    auto publish = [&](uint32_t slot) {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kTBarAFull + slot));
    };

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t pt = (int64_t)tile * kTileM + row;
      const bool valid = pt < p.n_points;
      float x[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
      if (valid && !p.x_embedded) {
        const int64_t ray = pt / p.samples_per_ray;
        if (p.rays_o) {
          // sample point built here: pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
          // (utils/create_data.py:486-487, :517): one rounded product and one rounded sum, as torch evaluates it
          const float z = __ldg(p.z_vals + pt);
#pragma unroll
          for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(__ldg(p.rays_o + ray * 3 + c), __fmul_rn(__ldg(p.rays_d + ray * 3 + c), z));
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) x[c] = __ldg(p.pts + pt * 3 + c);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) dir[c] = __ldg(p.viewdirs + ray * 3 + c);
      }
      float enc[32];
      // embedded-input mode (NeRF.forward called directly with the [P,90] tensor of run_network :66-72): gather the
      // 63 point features / 27 view features in the kernel's slot order instead of evaluating sin/cos
      auto gather = [&](int nf, int base) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int f = teacher_slot_to_feature(32 * (int)hf + i, nf);
          enc[i] = (valid && f >= 0) ? __ldg(p.x_embedded + pt * 90 + base + f) : 0.f;
        }
      };
      if (p.x_embedded) gather(10, 0); else encode_generic<10>(x, hf, enc);
      t_store_a_half(smem_base + kTSmemA, row, hf, enc);           // T0 input: slot 0
      publish(0);

      float alpha_part = 0.f, rgb_part[3] = {0.f, 0.f, 0.f};
      for (int l = 0; l < kTeacherLayers; ++l) {
        mbar_wait(bar(kTBarAccFull), acc_phase);
        acc_phase ^= 1u;
        tc_fence_after_sync();
        const float* bias = biases + l * kWidth;
        const bool relu = l != 8;
        const int ncol_chunks = l == 9 ? 2 : 4;                    // the view branch is 128 wide
        for (int c = 0; c < ncol_chunks; ++c) {
          const uint32_t col = 64u * c + 32u * hf;
          uint32_t r[32];
          tmem_ld32(tmem_row + 256u * (l & 1) + col, r);
          float4 bq[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) bq[i] = __ldg(reinterpret_cast<const float4*>(bias + col) + i);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[4 * i + 0] = fmaf(__uint_as_float(r[4 * i + 0]), kInvWeightScale, bq[i].x);   // exact power-of-two rescale
            v[4 * i + 1] = fmaf(__uint_as_float(r[4 * i + 1]), kInvWeightScale, bq[i].y);
            v[4 * i + 2] = fmaf(__uint_as_float(r[4 * i + 2]), kInvWeightScale, bq[i].z);
            v[4 * i + 3] = fmaf(__uint_as_float(r[4 * i + 3]), kInvWeightScale, bq[i].w);
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (l == 7) {   // alpha = alpha_linear(h) on the fp32 activations (:390)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(alpha_w + col) + i);
              alpha_part = fmaf(v[4 * i], w.x, alpha_part); alpha_part = fmaf(v[4 * i + 1], w.y, alpha_part);
              alpha_part = fmaf(v[4 * i + 2], w.z, alpha_part); alpha_part = fmaf(v[4 * i + 3], w.w, alpha_part);
            }
          }
          if (l < 9) {
            t_store_a_half(smem_base + kTSmemA + c * kAChunkBytes, row, hf, v);
            publish(c);
          } else {       // rgb = rgb_linear(relu(views_linear(...))) (:395-398)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(rgb_w + ch * 128 + col) + i);
                rgb_part[ch] = fmaf(v[4 * i], w.x, rgb_part[ch]); rgb_part[ch] = fmaf(v[4 * i + 1], w.y, rgb_part[ch]);
                rgb_part[ch] = fmaf(v[4 * i + 2], w.z, rgb_part[ch]); rgb_part[ch] = fmaf(v[4 * i + 3], w.w, rgb_part[ch]);
              }
            }
          }
        }
        if (l == 4 || l == 8) {
          // the 5th K-chunk of the next layer goes into slot 0 once its first chunk has been consumed:
          //   after T4: the point embedding again (skip connection, :383-384); after T8: the view-direction embedding (:392)
          if (p.x_embedded) { if (l == 4) gather(10, 0); else gather(4, 63); }
          else if (l == 4) encode_generic<10>(x, hf, enc);
          else encode_generic<4>(dir, hf, enc);
          mbar_wait(bar(kTBarAEmpty), l == 4 ? 0u : 1u);
          t_store_a_half(smem_base + kTSmemA, row, hf, enc);
          publish(0);
        }
      }
      // combine the column halves, write raw = [rgb, alpha] (:399)
      tc_fence_before_sync();
      if (hf == 1) {
        out_smem[row * 4 + 0] = rgb_part[0]; out_smem[row * 4 + 1] = rgb_part[1];
        out_smem[row * 4 + 2] = rgb_part[2]; out_smem[row * 4 + 3] = alpha_part;
      }
      named_bar_sync(1, kTEpiWarps * 32);
      if (hf == 0 && valid) {
        float4 o;
        o.x = rgb_part[0] + out_smem[row * 4 + 0] + __ldg(scal + 1);
        o.y = rgb_part[1] + out_smem[row * 4 + 1] + __ldg(scal + 2);
        o.z = rgb_part[2] + out_smem[row * 4 + 2] + __ldg(scal + 3);
        o.w = alpha_part + out_smem[row * 4 + 3] + __ldg(scal + 0);
        reinterpret_cast<float4*>(p.raw)[pt] = o;
      }
      named_bar_sync(1, kTEpiWarps * 32);   // out_smem is reused by the next tile
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ----------------------------------------------------------------------------------------------
// packing: teacher state_dict (flat, order pts_linears.0..7, views_linears.0, feature_linear, alpha_linear,
// rgb_linear; weight then bias each; NeRF.__init__ :357-375) -> 39 image pairs + fp32 tables
// ----------------------------------------------------------------------------------------------
struct TeacherOffsets {
  int64_t w[12], b[12];
};
__host__ __device__ inline TeacherOffsets teacher_offsets() {
  // linears in state_dict order: 0..7 pts_linears, 8 views_linears.0, 9 feature_linear, 10 alpha_linear, 11 rgb_linear
  const int out_dim[12] = {256, 256, 256, 256, 256, 256, 256, 256, 128, 256, 1, 3};
  const int in_dim[12] = {63, 256, 256, 256, 256, 319, 256, 256, 283, 256, 256, 128};
  TeacherOffsets t;
  int64_t off = 0;
  for (int i = 0; i < 12; ++i) {
    t.w[i] = off; off += (int64_t)out_dim[i] * in_dim[i];
    t.b[i] = off; off += out_dim[i];
  }
  return t;
}

__global__ void __launch_bounds__(256) teacher_pack_images_kernel(const float* __restrict__ params, uint8_t* __restrict__ packed) {
  const int ip = blockIdx.y;                               // image pair 0..38
  const int unit = blockIdx.x * blockDim.x + threadIdx.x;  // 0..2047
  const int n = unit >> 3, j = unit & 7;
  const TeacherOffsets t = teacher_offsets();
  // which layer / chunk
  const int chunks[kTeacherLayers] = {1, 4, 4, 4, 4, 5, 4, 4, 4, 5};
  int l = 0, c = ip;
  while (c >= chunks[l]) { c -= chunks[l]; ++l; }
  // kernel layer -> state_dict linear: T0..T7 = pts_linears, T8 = feature_linear (9), T9 = views_linears.0 (8)
  const int lin = l < 8 ? l : (l == 8 ? 9 : 8);
  const int in_dim = l == 0 ? 63 : (l == 5 ? 319 : (l == 9 ? 283 : 256));
  const int out_dim = l == 9 ? 128 : 256;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int slot = 8 * j + e;
    int col = -1;   // column of the reference weight matrix
    if (l == 0) col = teacher_slot_to_feature(slot, 10);
    else if (l == 5) col = c < 4 ? 63 + 64 * c + slot : teacher_slot_to_feature(slot, 10);   // cat([input_pts, h]) :384
    else if (l == 9) col = c < 4 ? 64 * c + slot : (teacher_slot_to_feature(slot, 4) >= 0 ? 256 + teacher_slot_to_feature(slot, 4) : -1);  // cat([feature, views]) :392
    else col = 64 * c + slot;
    v[e] = (col >= 0 && n < out_dim) ? params[t.w[lin] + (int64_t)n * in_dim + col] : 0.f;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split2(v[2 * e] * kWeightScale, v[2 * e + 1] * kWeightScale, hi[e], lo[e]);   // see ptx.cuh
  const uint32_t off = sw128_offset(n, 8 * j);
  uint8_t* img = packed + (int64_t)(2 * ip) * kWImageBytes;
  *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(img + kWImageBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(256) teacher_pack_tables_kernel(const float* __restrict__ params, uint8_t* __restrict__ packed) {
  const int col = threadIdx.x;
  const TeacherOffsets t = teacher_offsets();
  float* tab = reinterpret_cast<float*>(packed + kTeacherPackOffTables);
  for (int l = 0; l < 8; ++l) tab[l * kWidth + col] = params[t.b[l] + col];
  tab[8 * kWidth + col] = params[t.b[9] + col];                                 // feature_linear bias
  tab[9 * kWidth + col] = col < 128 ? params[t.b[8] + col] : 0.f;               // views_linears.0 bias
  float* alpha_w = tab + kTeacherLayers * kWidth;
  alpha_w[col] = params[t.w[10] + col];
  float* rgb_w = alpha_w + kWidth;
  for (int i = col; i < 3 * 128; i += 256) rgb_w[i] = params[t.w[11] + i];
  float* scal = rgb_w + 3 * 128;
  if (col == 0) scal[0] = params[t.b[10]];
  if (col >= 1 && col <= 3) scal[col] = params[t.b[11] + col - 1];
}

cudaError_t launch_teacher_pack(const float* params, void* packed, cudaStream_t stream) {
  dim3 grid(2048 / 256, kTeacherImagePairs);
  teacher_pack_images_kernel<<<grid, 256, 0, stream>>>(params, static_cast<uint8_t*>(packed));
  teacher_pack_tables_kernel<<<1, 256, 0, stream>>>(params, static_cast<uint8_t*>(packed));
  return cudaGetLastError();
}

cudaError_t launch_teacher(const TeacherParams& p, int grid, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(r2l_teacher_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmemBytes);
  if (e != cudaSuccess) return e;
  r2l_teacher_kernel<<<grid, kTThreads, kTSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace r2l
