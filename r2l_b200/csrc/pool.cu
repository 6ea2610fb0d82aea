// Hard-example ray pool on the device, two launches per iteration, both capturable in a CUDA graph.
//
// Reference: /root/reference/main.py
//   :1325-1347  draw   - once the pool is full, n_hard_out pool rays at np.random.permutation(pool)[:n_hard_out] are appended
//                        to the fresh batch;
//   :1410-1425  update - torch.sort of the per-ray mean squared error of the FRESH rays; the n_hard_in rays with the largest
//                        error are appended to the pool (filling) or overwrite the first n_hard_in drawn slots (full).
// The reference does both on the host between iterations (a device->host sync each); here
//   r2l_pool_draw_kernel    slot_j = P(j), j < n_out, with P a keyed pseudo-random PERMUTATION of [0, size) evaluated per
//                           element (cycle-walking Feistel network on the next even power of two, keyed by the seed and the
//                           device-resident iteration counter): n_out distinct slots, no sort, no scan; the thread copies row
//                           slot_j into the batch;
//   r2l_pool_update_kernel  one CTA: 4-pass radix select of the k-th largest error (keys = the float bits, errors are >= 0),
//                           ordered compaction of the selected ray indices, row scatter into the pool.
// HBM-bound integer/byte work: 4 B/ray read four-and-a-bit times from L2 (the loss kernel has just written the errors)
// and 72 B per moved row.
#include "kernels.cuh"

namespace r2l {

__host__ __device__ inline uint32_t mix32(uint32_t h) {     // murmur3 finaliser
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}

// bijection of [0, 2^(2 half_bits)): 6 Feistel rounds
__host__ __device__ inline uint32_t feistel(uint32_t x, int half_bits, uint32_t k0, uint32_t k1) {
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t l = x >> half_bits, r = x & mask;
  for (int round = 0; round < 6; ++round) {
    const uint32_t f = mix32(r * 0x9e3779b1u + k0 + (uint32_t)round * 0x7f4a7c15u) ^ mix32(k1 + (uint32_t)round);
    const uint32_t nl = r;
    r = (l ^ f) & mask;
    l = nl;
  }
  return (l << half_bits) | r;
}

// slot j of the permutation of [0, size) keyed by (seed, step); j < size.  Same code on the host (r2l_pool_slot_host: tests).
__host__ __device__ inline uint32_t pool_slot(uint32_t j, uint32_t size, uint64_t seed, uint64_t step) {
  int bits = 1;
  while (bits < 32 && (1ull << bits) < (unsigned long long)size) ++bits;
  const int half_bits = (bits + 1) >> 1;
  const uint32_t k0 = mix32((uint32_t)seed ^ (uint32_t)step);
  const uint32_t k1 = mix32((uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32) ^ 0x5bd1e995u);
  uint32_t x = j;
  do { x = feistel(x, half_bits, k0, k1); } while (x >= size);      // cycle walking: stays a permutation of [0, size)
  return x;
}

__global__ void __launch_bounds__(256) r2l_pool_draw_kernel(const float* __restrict__ pool_rows, const int* __restrict__ state,
                                                            int n_out, uint64_t seed,
                                                            const long long* __restrict__ counters, float* __restrict__ dst_rows,
                                                            int* __restrict__ slots_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const uint32_t size = (uint32_t)state[0];
  if (size == 0u || (uint32_t)j >= size) {      // cannot happen through the host API (n_out <= size); keep the row defined
    slots_out[j] = 0;
    for (int c = 0; c < 9; ++c) dst_rows[(int64_t)j * 9 + c] = 0.f;
    return;
  }
  const uint32_t x = pool_slot((uint32_t)j, size, seed, (uint64_t)counters[0]);
  slots_out[j] = (int)x;
  const float* src = pool_rows + (int64_t)x * 9;
  float* dst = dst_rows + (int64_t)j * 9;
#pragma unroll
  for (int c = 0; c < 9; ++c) dst[c] = __ldg(src + c);
}

constexpr int kPoolThreads = 1024;

// exclusive scan of `v` over the block (kPoolThreads threads); returns the exclusive prefix, `total` = block sum
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  __syncthreads();                       // warp_sums may still be read from the previous call
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += o;
    }
    warp_sums[32 + lane] = w;            // inclusive over warps
  }
  __syncthreads();
  total = warp_sums[32 + 31];
  return incl - v + (warp > 0 ? warp_sums[32 + warp - 1] : 0);
}

__global__ void __launch_bounds__(kPoolThreads, 1) r2l_pool_update_kernel(const float* __restrict__ rays9, const float* __restrict__ err,
                                                                          int n, int k, float* __restrict__ pool_rows,
                                                                          int* __restrict__ state, const int* __restrict__ slots_out,
                                                                          int* __restrict__ picked) {
  __shared__ int hist[256];
  __shared__ int warp_sums[64];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining;
  const int tid = threadIdx.x;
  const uint32_t* keys = reinterpret_cast<const uint32_t*>(err);
  // canonical key: errors are >= 0 so the float bits order like the values; -0.0 -> +0.0; NaN sorts last (= largest), as in torch.sort
  auto key_of = [&](int i) { const uint32_t b = keys[i]; return b == 0x80000000u ? 0u : b; };

  if (tid == 0) { s_prefix = 0u; s_remaining = k; }
  uint32_t mask = 0u;
  for (int pass = 3; pass >= 0; --pass) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const int shift = 8 * pass;
    for (int i = tid; i < n; i += kPoolThreads) {
      const uint32_t key = key_of(i);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (tid < 32) {
      // warp 0 finds the digit whose bin holds the remaining-th largest key: lane l owns bins 8l .. 8l+7, a suffix sum over
      // the lanes gives the keys in higher bins (a serial walk over 256 shared-memory words by one thread cost 4 us per pass)
      int c[8], mine = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[8 * tid + j]; mine += c[j]; }
      int incl = mine;                     // keys in the bins of lanes >= this one
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_down_sync(0xffffffffu, incl, d);
        if (tid + d < 32) incl += o;
      }
      const int above = incl - mine, remaining = s_remaining;
      __syncwarp();
      if (above < remaining && remaining <= above + mine) {       // exactly one lane (k <= n keeps remaining <= matching keys)
        int rem = remaining - above, b = 0;
        bool found = false;
#pragma unroll
        for (int j = 7; j >= 1; --j) {     // walk down from the lane's largest digit
          if (!found) {
            if (c[j] >= rem) { b = j; found = true; }
            else rem -= c[j];
          }
        }
        s_remaining = rem;                 // how many keys with this digit (and the prefix so far) are still to be taken
        s_prefix = prefix | ((uint32_t)(8 * tid + b) << shift);
      }
    }
    mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t thr = s_prefix;           // the k-th largest key
  const int take_eq = s_remaining;         // keys equal to it that are selected (the lowest indices)
  const int n_gt = k - take_eq;            // keys above it
  const bool append = slots_out == nullptr;
  const int base = append ? state[0] : 0;
  int run_gt = 0, run_eq = 0;
  for (int start = 0; start < n; start += kPoolThreads) {
    const int i = start + tid;
    uint32_t key = 0u;
    bool gt = false, eq = false;
    if (i < n) {
      key = key_of(i);
      gt = key > thr;
      eq = key == thr;
    }
    int total = 0;
    const int packed_excl = block_excl_scan((gt ? 1 : 0) | (eq ? (1 << 16) : 0), warp_sums, total);
    const int my_gt = run_gt + (packed_excl & 0xffff), my_eq = run_eq + (packed_excl >> 16);
    int j = -1;
    if (gt) j = my_gt;
    else if (eq && my_eq < take_eq) j = n_gt + my_eq;
    if (j >= 0) {
      const int64_t dst_row = append ? (int64_t)base + j : (int64_t)slots_out[j];
      const float* src = rays9 + (int64_t)i * 9;
      float* dst = pool_rows + dst_row * 9;
#pragma unroll
      for (int c = 0; c < 9; ++c) dst[c] = src[c];
      if (picked) picked[j] = i;
    }
    run_gt += total & 0xffff;
    run_eq += total >> 16;
  }
  if (append && tid == 0) state[0] = base + k;
}

uint32_t pool_slot_host(uint32_t j, uint32_t size, uint64_t seed, uint64_t step) { return pool_slot(j, size, seed, step); }

cudaError_t launch_pool_draw(const float* pool_rows, const int* state, int n_out, uint64_t seed, const long long* counters,
                             float* dst_rows, int* slots_out, cudaStream_t stream) {
  r2l_pool_draw_kernel<<<(n_out + 255) / 256, 256, 0, stream>>>(pool_rows, state, n_out, seed, counters, dst_rows, slots_out);
  return cudaGetLastError();
}

cudaError_t launch_pool_update(const float* rays9, const float* err, int n, int k, float* pool_rows, int* state, const int* slots_out,
                               int* picked, cudaStream_t stream) {
  r2l_pool_update_kernel<<<1, kPoolThreads, 0, stream>>>(rays9, err, n, k, pool_rows, state, slots_out, picked);
  return cudaGetLastError();
}

}  // namespace r2l
