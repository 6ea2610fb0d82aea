// Weight-gradient kernel: dW_l[o,i] = sum_rays dY_l[ray,o] * X_l[ray,i] and db_l[o] = sum_rays dY_l[ray,o]
// for the head and the 86 body Linears, as tcgen05 GEMMs whose K dimension is the ray axis.
//
// Both operands are the bf16 hi/lo operand images the chain kernels stored (chain.cu): X_l = the A operand
// the forward pass fed to Linear l, dY_l = the A operand the backward pass built from dL/d(output of Linear l).
// An image chunk is [128 rays][64 features] with 128-byte swizzled rows, i.e. exactly a UMMA *MN-major*
// SWIZZLE_128B tile when the ray axis plays K — so the same bytes serve as K-major A operand in the chain
// kernels and as MN-major A/B operands here, with no transposition anywhere.
//
// One CTA per unit (86 body layers + 4 head column groups of 256 features); it loops over all ray tiles:
//   warp 0     producer: per 32-ray stage, 16 bulk copies of 4 KiB (dY: 4 chunks x {hi,lo}, X: same) -> 64 KiB stage
//   warp 1     MMA issuer: per stage 2 k-steps x 2 output halves x 3 split terms, M=128 N=256 K=16
//   warp 2     TMEM allocator: two fp32 accumulators [128 x 256] (output rows 0-127 and 128-255) = 512 columns
//   warps 4-7  bias gradient: column sums of dY straight from the staged smem tiles; then the epilogue
//              (TMEM -> registers -> global)
// Split units (DwParams::unit_splits[u] > 1): a unit is cut into ray-tile ranges, one CTA ("piece") each; a piece writes
// its partial dW / db to scratch and the LAST piece of the unit to finish (atomic ticket) sums the partials in fixed
// order - the result does not depend on which piece was last.  CTAs are ordered by the time their layer becomes
// available.  When the kernel overlaps the backward chain on otherwise idle SMs the host (c_api.cu: dw_schedule) leaves
// the units that are released early whole - they have the whole chain to finish in and cost no scratch traffic -
// and cuts the later ones ever finer, so that the pieces still running when the chain ends are short.
// Autograd equivalent in the reference: the weight/bias gradients torch.autograd produces for every
// nn.Linear of NeRF_v3_2 (model/nerf_raybased.py:500,:453-456) under loss.backward() (main.py:1404).
#include "kernels.cuh"
#include "ptx.cuh"

namespace r2l {

constexpr int kDwThreads = 256;
constexpr int kDwStages = 3;
constexpr int kDwRowsPerStage = 32;
constexpr uint32_t kDwPiece = kDwRowsPerStage * 128;   // 4096: 32 rays of one plane of one chunk
constexpr uint32_t kDwOperand = 8 * kDwPiece;          // 32768: 4 chunks x 2 planes
constexpr uint32_t kDwStageBytes = 2 * kDwOperand;     // dY + X
constexpr uint32_t kDwSmemBar = kDwStages * kDwStageBytes;  // 196608
constexpr uint32_t kDwSmemBytes = kDwSmemBar + 128 + 1024;
static_assert(kDwSmemBytes <= 232448, "exceeds the 227 KiB dynamic shared memory limit");
constexpr int kDwUnits = kBodyLayers + 4;

__global__ void __launch_bounds__(kDwThreads, 1) r2l_dw_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + kDwSmemBar;
  auto bar_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto bar_empty = [&](uint32_t s) { return bar0 + 8u * (kDwStages + s); };
  const uint32_t bar_done = bar0 + 8u * (2 * kDwStages);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kDwSmemBar + 8 * (2 * kDwStages + 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA order = order in which the backward sweep releases the layers: body layer 85 first ... layer 0, then the head
  int ob = 0;
  while (ob + 1 < kDwUnits && (int)blockIdx.x >= (int)p.unit_first[ob + 1]) ++ob;
  const int splits = p.unit_splits[ob], split = (int)blockIdx.x - (int)p.unit_first[ob];
  const bool is_head = ob >= kBodyLayers;
  const int layer = kBodyLayers - 1 - ob;  // body layer 0..85 when !is_head
  const int hg = ob - kBodyLayers;         // head feature group 0..3 (256 encoded features each)
  const int unit = is_head ? ob : layer;   // index used by the debug stamps
  const int tile_lo = (int)((int64_t)p.num_tiles * split / splits);
  const int tile_hi = (int)((int64_t)p.num_tiles * (split + 1) / splits);
  // chunk offsets inside a tile's saved images (see chain.cu for the order they are written in)
  const int x_chunk0 = is_head ? 4 * hg : kSamples + 4 * layer;
  const int dy_chunk0 = is_head ? kAChunks + 4 * (kBodyLayers - 1) : kAChunks + 4 * (kBodyLayers - 2 - layer);
  const int num_stages_total = (tile_hi - tile_lo) * (kTileM / kDwRowsPerStage);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1 + 4);   // MMA commit + the four bias warps
    }
    mbar_init(bar_done, 1);
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
      if (p.times) p.times[unit * 4 + 0] = global_timer_ns();
      if (p.ready != nullptr) {
        // concurrent mode: the backward chain kernel is still running on other SMs; wait until every tile has stored
        // the dY operand of this unit's layer (group index = position of that layer in the backward sweep)
        const int group = is_head ? kBodyLayers : ob;
        unsigned ns = 64;
        while (flag_acquire_load(p.ready + group) < p.ready_target) {
          __nanosleep(ns);
          if (ns < 2048) ns <<= 1;
        }
        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire -> async-proxy (TMA) reads
      }
      if (p.times) p.times[unit * 4 + 1] = global_timer_ns();
      for (int it = 0; it < num_stages_total; ++it) {
        const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1u;
        const int tile = tile_lo + (it >> 2), qr = it & 3;
        mbar_wait(bar_empty(s), ph ^ 1u);
        mbar_arrive_expect_tx(bar_full(s), kDwStageBytes);
        const uint8_t* dy = p.bwd_saved + ((int64_t)tile * kBwdSavedChunks + dy_chunk0) * kAChunkBytes + qr * kDwPiece;
        const uint8_t* x = p.fwd_saved + ((int64_t)tile * kFwdSavedChunks + x_chunk0) * kAChunkBytes + qr * kDwPiece;
        const uint32_t dst = smem_base + s * kDwStageBytes;
#pragma unroll
        for (int cp = 0; cp < 8; ++cp) {   // cp = chunk*2 + plane; planes are 16 KiB apart inside a 32 KiB chunk
          bulk_g2s(dst + cp * kDwPiece, dy + (int64_t)cp * kPlaneBytes, kDwPiece, bar_full(s));
          bulk_g2s(dst + kDwOperand + cp * kDwPiece, x + (int64_t)cp * kPlaneBytes, kDwPiece, bar_full(s));
        }
      }
    }
  } else if (warp == 1) {
    {   // the whole warp waits, one elected lane issues (keeps the descriptors in uniform registers, see chain.cu)
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 1, 1);   // both operands MN-major
      constexpr uint32_t kLbo = 2 * kDwPiece;   // next 64-feature group of the same plane
      constexpr uint32_t kSbo = 1024;           // next 8 rays
      for (int it = 0; it < num_stages_total; ++it) {
        const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1u;
        mbar_wait(bar_full(s), ph);
        tc_fence_after_sync();
        const uint32_t dy = smem_base + s * kDwStageBytes, x = dy + kDwOperand;
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t d = tmem_base + 256u * half;
              const uint32_t a_hi = dy + (4 * half) * kDwPiece + ks * 2048, a_lo = a_hi + kDwPiece;
              const uint32_t b_hi = x + ks * 2048, b_lo = b_hi + kDwPiece;
              const uint32_t first = (it == 0 && ks == 0) ? 0u : 1u;
              umma_bf16(d, umma_desc_sw128(a_hi, kLbo, kSbo), umma_desc_sw128(b_hi, kLbo, kSbo), idesc, first);
              umma_bf16(d, umma_desc_sw128(a_lo, kLbo, kSbo), umma_desc_sw128(b_hi, kLbo, kSbo), idesc, 1u);
              umma_bf16(d, umma_desc_sw128(a_hi, kLbo, kSbo), umma_desc_sw128(b_lo, kLbo, kSbo), idesc, 1u);
            }
          }
          umma_commit(bar_empty(s));
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(bar_done);
    }
  } else if (warp >= 4) {
    const uint32_t t = (warp - 4) * 32 + lane;   // 0..127: owns output columns 2t, 2t+1 of dY for the bias sum
    const uint32_t c = t >> 5;                   // chunk of those columns
    const uint32_t k = (2 * t) & 63;             // position inside the chunk
    float s0 = 0.f, s1 = 0.f;
    for (int it = 0; it < num_stages_total; ++it) {
      const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1u;
      mbar_wait(bar_full(s), ph);
      const uint8_t* dy = smem_gen + s * kDwStageBytes;
#pragma unroll 8
      for (uint32_t r = 0; r < kDwRowsPerStage; ++r) {
        const uint32_t off = r * 128u + ((((k >> 3) ^ (r & 7u)) << 4) | ((k & 7u) << 1));
        const uint32_t hi = *reinterpret_cast<const uint32_t*>(dy + (2 * c) * kDwPiece + off);
        const uint32_t lo = *reinterpret_cast<const uint32_t*>(dy + (2 * c + 1) * kDwPiece + off);
        s0 += __uint_as_float(hi << 16) + __uint_as_float(lo << 16);
        s1 += __uint_as_float(hi & 0xFFFF0000u) + __uint_as_float(lo & 0xFFFF0000u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty(s));
    }
    // where the results go: straight into the gradient buffer, or (split mode) into this piece's scratch slot
    constexpr int kUnitFloats = kWidth * kWidth + kWidth;            // dW [256][256] + db [256]
    float* part = splits > 1 ? p.partials + (int64_t)blockIdx.x * kUnitFloats : nullptr;
    auto head_feature = [&](int col) {
      const int chunk = 4 * hg + (col >> 6), slot = col & 63;
      return p.input_kind == kInputX ? (64 * chunk + slot < kInDim ? 64 * chunk + slot : -1) : fused_slot_to_feature(chunk, slot);
    };
    // bias gradient (head groups all compute the same sum; group 0 writes it)
    if (part) {
      part[kWidth * kWidth + 2 * t] = s0;
      part[kWidth * kWidth + 2 * t + 1] = s1;
    } else if (!is_head || hg == 0) {
      float* db = p.grads + (is_head ? kOffHeadB : off_body_b(layer)) + 2 * t;
      if (p.accumulate) { db[0] += s0; db[1] += s1; } else { db[0] = s0; db[1] = s1; }
    }
    // epilogue: both accumulators -> global
    mbar_wait(bar_done, 0);
    tc_fence_after_sync();
    const uint32_t q = warp & 3;
    const uint32_t tmem_row = tmem_base + ((q * 32u) << 16);
    for (int half = 0; half < 2; ++half) {
      const int o = half * 128 + q * 32 + lane;   // output feature (row of dW)
      for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t r[32];
        tmem_ld32(tmem_row + 256u * half + 32u * c8, r);
        tmem_ld_wait();
        if (part || !is_head) {
          float4* dst = reinterpret_cast<float4*>((part ? part : p.grads + off_body_w(layer)) + (int64_t)o * kWidth + 32 * c8);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 v = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                   __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
            if (!part && p.accumulate) { const float4 a = dst[i]; v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
            dst[i] = v;
          }
        } else {
          float* dst = p.grads + kOffHeadW + (int64_t)o * kInDim;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int feat = head_feature(32 * c8 + i);
            if (feat >= 0) {
              const float v = __uint_as_float(r[i]);
              if (p.accumulate) dst[feat] += v; else dst[feat] = v;
            }
          }
        }
      }
    }
    if (part) __threadfence();   // partial sums visible before this CTA takes its ticket (below, whole CTA)
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
  if (splits > 1) {
    // ticket: the last piece of this unit to finish reduces all partials in index order (deterministic result),
    // with the whole CTA and 8 independent float4 loads per thread in flight (the loop is L2-latency bound)
    constexpr int kUnitFloats = kWidth * kWidth + kWidth;
    int* ticket_smem = reinterpret_cast<int*>(smem_gen + kDwSmemBar + 96);
    if (threadIdx.x == 0) *ticket_smem = atomicAdd(p.tickets + ob, 1);
    __syncthreads();
    if (*ticket_smem == splits - 1) {
      __threadfence();
      auto head_feature = [&](int col) {
        const int chunk = 4 * hg + (col >> 6), slot = col & 63;
        return p.input_kind == kInputX ? (64 * chunk + slot < kInDim ? 64 * chunk + slot : -1) : fused_slot_to_feature(chunk, slot);
      };
      const float* base = p.partials + (int64_t)p.unit_first[ob] * kUnitFloats;
      constexpr int kVecs = kUnitFloats / 4;   // 16448
      constexpr int kU = 8;
      for (int v0 = (int)threadIdx.x; v0 < kVecs; v0 += kU * kDwThreads) {
        float4 acc[kU];
#pragma unroll
        for (int j = 0; j < kU; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sp = 0; sp < splits; ++sp) {
#pragma unroll
          for (int j = 0; j < kU; ++j) {
            const int v = v0 + kDwThreads * j;
            if (v < kVecs) {
              const float4 x = __ldcg(reinterpret_cast<const float4*>(base + (int64_t)sp * kUnitFloats) + v);
              acc[j].x += x.x; acc[j].y += x.y; acc[j].z += x.z; acc[j].w += x.w;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < kU; ++j) {
          const int v = v0 + kDwThreads * j;
          if (v >= kVecs) continue;
          const int idx = 4 * v;
          const float a[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
          if (idx >= kWidth * kWidth) {
            if (!is_head || hg == 0) {
              float* db = p.grads + (is_head ? kOffHeadB : off_body_b(layer)) + (idx - kWidth * kWidth);
#pragma unroll
              for (int e = 0; e < 4; ++e) db[e] = p.accumulate ? db[e] + a[e] : a[e];
            }
          } else if (!is_head) {
            float4* w = reinterpret_cast<float4*>(p.grads + off_body_w(layer) + idx);
            float4 o4 = acc[j];
            if (p.accumulate) { const float4 c4 = *w; o4.x += c4.x; o4.y += c4.y; o4.z += c4.z; o4.w += c4.w; }
            *w = o4;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int feat = head_feature((idx + e) & (kWidth - 1));
              if (feat >= 0) {
                float* w = p.grads + kOffHeadW + (int64_t)((idx + e) >> 8) * kInDim + feat;
                *w = p.accumulate ? *w + a[e] : a[e];
              }
            }
          }
        }
      }
    }
  }
  if (p.times && threadIdx.x == 0) p.times[unit * 4 + 3] = global_timer_ns();
}

// tail.0.weight / tail.0.bias gradients: dW_t[c,j] = sum_n dlogit[n,c] (z_43 + h)[n,j]; 768+3 outputs, CUDA cores.
__global__ void __launch_bounds__(256) r2l_tail_grad_kernel(const __grid_constant__ TailGradParams p) {
  __shared__ float dl[128][3];
  const int64_t n0 = (int64_t)blockIdx.x * 128;
  if (threadIdx.x < 128) {
    const int64_t n = n0 + threadIdx.x;
    float v[3] = {0.f, 0.f, 0.f};
    if (n < p.n_rays) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float y = p.rgb[n * 3 + c];
        v[c] = p.grad_rgb[n * 3 + c] * y * (1.f - y);
      }
    }
    dl[threadIdx.x][0] = v[0]; dl[threadIdx.x][1] = v[1]; dl[threadIdx.x][2] = v[2];
  }
  __syncthreads();
  const int j = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  const int rows = (int)((p.n_rays - n0) < 128 ? (p.n_rays - n0) : 128);
  for (int r = 0; r < rows; ++r) {
    const float z = p.zf[(n0 + r) * kWidth + j];
    a0 = fmaf(dl[r][0], z, a0); a1 = fmaf(dl[r][1], z, a1); a2 = fmaf(dl[r][2], z, a2);
  }
  atomicAdd(p.grads + kOffTailW + j, a0);
  atomicAdd(p.grads + kOffTailW + kWidth + j, a1);
  atomicAdd(p.grads + kOffTailW + 2 * kWidth + j, a2);
  if (j < 3) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += dl[r][j];
    atomicAdd(p.grads + kOffTailB + j, s);
  }
}

cudaError_t launch_dw(const DwParams& p, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(r2l_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDwSmemBytes);
  if (e != cudaSuccess) return e;
  r2l_dw_kernel<<<p.num_ctas, kDwThreads, kDwSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_tail_grads(const TailGradParams& p, bool zero_first, cudaStream_t stream) {
  if (zero_first) {
    cudaError_t e = cudaMemsetAsync(p.grads + kOffTailW, 0, (kOutDim * kWidth + kOutDim) * sizeof(float), stream);
    if (e != cudaSuccess) return e;
  }
  const int blocks = (int)((p.n_rays + 127) / 128);
  r2l_tail_grad_kernel<<<blocks, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace r2l
