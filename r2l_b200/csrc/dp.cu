// Data-parallel parameter update over NVLink / NVSwitch peer memory: ONE kernel per iteration does what the reference's
// nn.DataParallel + torch.optim.Adam do with a gradient reduce to GPU 0, an optimizer step there and a re-broadcast of all
// 23.7 MB of parameters before the next forward (main.py:37-42, :472-479, :1403-1406):
//
//   every rank (one process per GPU) owns a 1/world slice of the flat parameter buffer.  It reads THAT slice of the
//   gradient from all ranks' buffers through peer loads (reduce-scatter, summed in rank order), applies Adam to the
//   slice with its own slice of the moments, and writes the new parameters into every rank's parameter buffer through
//   peer stores (all-gather).  Two flag barriers in peer memory bracket it: "my gradient is complete" / "my slice is
//   written everywhere".
//
// Against all-reduce + Adam (NCCL: 143 us for the 23.7 MB buffer on 8 GPUs + 23 us Adam): each GPU moves 7/8 of the buffer in
// and 7/8 out over NVLink concurrently and updates 1/8 of the parameters.  Every parameter is computed exactly once, so the
// ranks' parameters are bit-identical by construction, and the sum order is fixed (rank 0 .. world-1): bit-reproducible.
// The buffers are plain cudaMalloc allocations shared through CUDA IPC handles (host side: c_api.cu, r2l_dp_*).
#include "kernels.cuh"
#include "ptx.cuh"

namespace r2l {

__device__ __forceinline__ void st_release_sys_u32(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
// peer data: system-scope relaxed loads (never served from a stale L1 line of an earlier iteration)
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* addr) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f1(const float* addr) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(addr) : "memory");
  return v;
}

// wait until every rank's flag in my flag block has reached `epoch`; a peer that never arrives (crashed process) must end in
// a trap, not in a hung GPU
__device__ __forceinline__ void wait_flags(const uint32_t* my_flags, int world, uint32_t epoch) {
  if (threadIdx.x < (unsigned)world) {
    const long long t0 = global_timer_ns();
    unsigned ns = 20;
    while ((int32_t)(ld_acquire_sys_u32(my_flags + threadIdx.x) - epoch) < 0) {
      __nanosleep(ns);
      if (ns < 640) ns <<= 1;
      if (global_timer_ns() - t0 > 20000000000ll) __trap();
    }
  }
  __syncthreads();
}

// WORLD > 0: the rank count as a compile-time constant, so that the peer loads of a thread (U positions x WORLD ranks = 16
// float4 in flight) are all issued before the first add - a thread that waits for one NVLink round trip per load moves
// ~400 GB/s per GPU, far below the links.  WORLD = 0: any rank count, one position at a time.
template <int WORLD>
__global__ void __launch_bounds__(256) r2l_dp_adam_kernel(const __grid_constant__ DpParams p) {
  constexpr int U = WORLD > 0 ? (16 / WORLD > 0 ? 16 / WORLD : 1) : 1;
  const int world = WORLD > 0 ? WORLD : p.world;
  __shared__ uint32_t s_epoch;
  __shared__ int s_last;
  if (threadIdx.x == 0) s_epoch = p.state[0] + 1u;   // state[0] = epochs completed (advanced by the last CTA below)
  __syncthreads();
  const uint32_t epoch = s_epoch;
  uint32_t* my_flags = p.flags[p.rank] + p.flag_base;   // one group of 32 flag words per concurrent launch slot

  // ---- barrier 1: every rank's gradient buffer is complete (this kernel runs behind the backward on its stream) ----
  if (blockIdx.x == 0 && threadIdx.x < (unsigned)world) {
    __threadfence_system();
    st_release_sys_u32(p.flags[threadIdx.x] + p.flag_base + p.rank, epoch);
  }
  wait_flags(my_flags, world, epoch);

  // ---- my slice: sum of the ranks' gradients (rank order), Adam, new parameters to every rank ----
  const float step_size = __ldg(p.hyper), inv_bc2_sqrt = __ldg(p.hyper + 1);
  const int64_t lo = p.shard_lo, hi = p.shard_hi;
  const int64_t n4 = (hi - lo) >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 g[U];
    if constexpr (WORLD > 0) {
      float4 x[U][WORLD];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
#pragma unroll
        for (int s = 0; s < WORLD; ++s)
          x[u][s] = (i < n4 && (s == p.rank || !(p.variant & 4)))
                        ? ((p.variant & 1) ? __ldcg(reinterpret_cast<const float4*>(p.grads[s] + lo + 4 * i)) : ld_relaxed_sys_f4(p.grads[s] + lo + 4 * i))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        g[u] = x[u][0];
#pragma unroll
        for (int s = 1; s < WORLD; ++s) { g[u].x += x[u][s].x; g[u].y += x[u][s].y; g[u].z += x[u][s].z; g[u].w += x[u][s].w; }
      }
    } else {
      g[0] = ld_relaxed_sys_f4(p.grads[0] + lo + 4 * i0);
      for (int s = 1; s < world; ++s) {
        const float4 x = ld_relaxed_sys_f4(p.grads[s] + lo + 4 * i0);
        g[0].x += x.x; g[0].y += x.y; g[0].z += x.z; g[0].w += x.w;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n4) break;
      const int64_t off = lo + 4 * i;
      float4 pp = *reinterpret_cast<const float4*>(p.params[p.rank] + off);
      float4 mm = reinterpret_cast<float4*>(p.exp_avg + off)[0];
      float4 vv = reinterpret_cast<float4*>(p.exp_avg_sq + off)[0];
      float* pa = &pp.x; const float* ga = &g[u].x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {   // the arithmetic of r2l_adam_kernel (optim.cu), i.e. of torch.optim.Adam
        ma[k] = ma[k] + (ga[k] - ma[k]) * p.w1;
        va[k] = __fmaf_rn(p.w2 * ga[k], ga[k], va[k] * p.beta2);
        const float denom = sqrtf(va[k]) * inv_bc2_sqrt + p.eps;
        pa[k] = pa[k] - step_size * (ma[k] / denom);
      }
      reinterpret_cast<float4*>(p.exp_avg + off)[0] = mm;
      reinterpret_cast<float4*>(p.exp_avg_sq + off)[0] = vv;
#pragma unroll
      for (int s = 0; s < (WORLD > 0 ? WORLD : 1); ++s)
        if (WORLD > 0 && (s == p.rank || !(p.variant & 2))) *reinterpret_cast<float4*>(p.params[s] + off) = pp;
      if (WORLD == 0)
        for (int s = 0; s < world; ++s) *reinterpret_cast<float4*>(p.params[s] + off) = pp;
    }
  }
  // tail of the slice (the buffer's length is not a multiple of 4: only the last rank's slice has one)
  for (int64_t off = lo + 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; off < hi; off += stride) {
    float g = 0.f;
    for (int s = 0; s < world; ++s) g += ld_relaxed_sys_f1(p.grads[s] + off);
    const float mi = p.exp_avg[off] + (g - p.exp_avg[off]) * p.w1;
    const float vi = __fmaf_rn(p.w2 * g, g, p.exp_avg_sq[off] * p.beta2);
    p.exp_avg[off] = mi; p.exp_avg_sq[off] = vi;
    const float pn = p.params[p.rank][off] - step_size * (mi / (sqrtf(vi) * inv_bc2_sqrt + p.eps));
    for (int s = 0; s < world; ++s) p.params[s][off] = pn;
  }

  // ---- barrier 2: the last CTA of this rank announces "my slice is written everywhere" and waits for everyone's ----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(reinterpret_cast<int*>(p.state + 1), 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < (unsigned)world) st_release_sys_u32(p.flags[threadIdx.x] + p.flag_base + 16 + p.rank, epoch);
  wait_flags(my_flags + 16, world, epoch);
  if (threadIdx.x == 0) {
    p.state[1] = 0;
    p.state[0] = epoch;
  }
}

cudaError_t launch_dp_adam(const DpParams& p, int grid, cudaStream_t stream) {
  switch (p.world) {
    case 2: r2l_dp_adam_kernel<2><<<grid, 256, 0, stream>>>(p); break;
    case 4: r2l_dp_adam_kernel<4><<<grid, 256, 0, stream>>>(p); break;
    case 8: r2l_dp_adam_kernel<8><<<grid, 256, 0, stream>>>(p); break;
    default: r2l_dp_adam_kernel<0><<<grid, 256, 0, stream>>>(p); break;
  }
  return cudaGetLastError();
}

}  // namespace r2l
