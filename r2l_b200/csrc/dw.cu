// Weight-gradient kernel: dW_l[o,i] = sum_rays dY_l[ray,o] * X_l[ray,i] and db_l[o] = sum_rays dY_l[ray,o]
// for the head and the 86 body Linears, as tcgen05 GEMMs whose K dimension is the ray axis.
//
// Both operands are the fp16 hi/lo operand images the chain kernels stored (chain.cu): X_l = the A operand
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
constexpr uint32_t kDwSmemStage = kDwSmemBar + 128;              // epilogue transposition: 4 warps x [32 rows][36 floats]
constexpr uint32_t kDwStageRow = 36;                              // floats per staged row (16-byte aligned, conflict-free)
constexpr uint32_t kDwSmemBytes = kDwSmemStage + 4 * 32 * kDwStageRow * 4 + 1024;
static_assert(kDwSmemBytes <= 232448, "exceeds the 227 KiB dynamic shared memory limit");
constexpr int kDwUnits = kBodyLayers + 4;

// smem after the stage ring: mbarriers (8 B each) full[3] empty[3] done tmem_free item_full[2] item_empty[2] = 12,
// then the TMEM address, the two-entry item ring and the ticket broadcast word
constexpr uint32_t kDwBarDone = 2 * kDwStages, kDwBarTmemFree = kDwBarDone + 1, kDwBarItemFull = kDwBarTmemFree + 1,
                   kDwBarItemEmpty = kDwBarItemFull + 2, kDwNumBars = kDwBarItemEmpty + 2;
static_assert(8 * kDwNumBars + 16 <= 128, "barrier block");

// Persistent: gridDim.x CTAs (at most one per SM) claim work items in order from the counter p.queue.  Item i is piece
// (i - unit_first[u]) of unit u; items are numbered in the order the backward chain releases their layers, so a CTA only
// ever waits on a readiness flag when nothing claimable is released yet, and CTAs that get their SM late (those of the
// chain kernel's SMs) simply join the queue.  Inside the CTA the producer lane runs ahead of the consumers by one item
// (two-entry item ring) so the operand stream does not drain between items; the accumulators are handed back to the MMA
// warp as soon as the epilogue has read them.
__global__ void __launch_bounds__(kDwThreads, 1) r2l_dw_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + kDwSmemBar;
  auto bar = [&](uint32_t i) { return bar0 + 8u * i; };
  auto bar_full = [&](uint32_t s) { return bar(s); };
  auto bar_empty = [&](uint32_t s) { return bar(kDwStages + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kDwSmemBar + 8 * kDwNumBars);
  volatile int* item_ring = reinterpret_cast<volatile int*>(smem_gen + kDwSmemBar + 8 * kDwNumBars + 4);   // [2]
  volatile int* ticket_smem = reinterpret_cast<volatile int*>(smem_gen + kDwSmemBar + 8 * kDwNumBars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1 + 4);   // MMA commit + the four bias warps
    }
    mbar_init(bar(kDwBarDone), 1);
    mbar_init(bar(kDwBarTmemFree), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(kDwBarItemFull + i), 1);
      mbar_init(bar(kDwBarItemEmpty + i), 1 + 4);   // MMA warp + the four epilogue warps
    }
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

  // what an item is: unit (release order: body layer 85 .. 0, then the four head column groups) and ray-tile range
  struct Item {
    int ob, split, splits, layer, hg, tile_lo, stages, x_chunk0, dy_chunk0;
    bool is_head;
  };
  auto decode = [&](int item) {
    Item w;
    int ob = 0;
    while (ob + 1 < kDwUnits && item >= (int)p.unit_first[ob + 1]) ++ob;
    w.ob = ob;
    w.splits = p.unit_splits[ob];
    w.split = item - (int)p.unit_first[ob];
    w.is_head = ob >= kBodyLayers;
    w.layer = kBodyLayers - 1 - ob;   // body layer 0..85 when !is_head
    w.hg = ob - kBodyLayers;          // head feature group 0..3 (256 encoded features each)
    w.tile_lo = (int)((int64_t)p.num_tiles * w.split / w.splits);
    const int tile_hi = (int)((int64_t)p.num_tiles * (w.split + 1) / w.splits);
    w.stages = (tile_hi - w.tile_lo) * (kTileM / kDwRowsPerStage);
    // chunk offsets inside a tile's saved images (see chain.cu for the order they are written in)
    w.x_chunk0 = w.is_head ? 4 * w.hg : kSamples + 4 * w.layer;
    w.dy_chunk0 = w.is_head ? kAChunks + 4 * (kBodyLayers - 1) : kAChunks + 4 * (kBodyLayers - 2 - w.layer);
    return w;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t git = 0;   // stage counter over all items of this CTA (the stage ring never restarts)
      for (uint32_t k = 0;; ++k) {
        mbar_wait(bar(kDwBarItemEmpty + (k & 1u)), ((k >> 1) & 1u) ^ 1u);
        int item = p.item_lo + atomicAdd(p.queue, 1);
        if (item >= p.item_hi) item = -1;
        item_ring[k & 1u] = item;
        mbar_arrive(bar(kDwBarItemFull + (k & 1u)));   // release: the ring entry is visible to whoever sees the phase
        if (item < 0) break;
        const Item w = decode(item);
        if (p.times) p.times[w.ob * 4 + 0] = global_timer_ns();
        if (p.ready != nullptr) {
          // concurrent mode: the backward chain kernel is still running on other SMs; wait until every tile has stored
          // the dY operand of this unit's layer (group index = position of that layer in the backward sweep)
          const int group = w.is_head ? kBodyLayers : w.ob;
          unsigned ns = 32;
          const long long t_wait0 = global_timer_ns();
          while (flag_acquire_load(p.ready + group) < p.ready_target) {
            __nanosleep(ns);
            if (ns < 1024) ns <<= 1;
            // the chain kernel this launch runs beside never became resident (SMs taken by another process / MPS client):
            // fail loudly instead of hanging the device (the GPU is expected to be exclusive)
            if (global_timer_ns() - t_wait0 > 4000000000ll) __trap();
          }
          asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire -> async-proxy (TMA) reads
        }
        if (p.times) p.times[w.ob * 4 + 1] = global_timer_ns();
        for (int it = 0; it < w.stages; ++it, ++git) {
          const uint32_t s = git % kDwStages, ph = (git / kDwStages) & 1u;
          const int tile = w.tile_lo + (it >> 2), qr = it & 3;
          mbar_wait(bar_empty(s), ph ^ 1u);
          mbar_arrive_expect_tx(bar_full(s), kDwStageBytes);
          const uint8_t* dy = p.bwd_saved + ((int64_t)tile * kBwdSavedChunks + w.dy_chunk0) * kAChunkBytes + qr * kDwPiece;
          const uint8_t* x = p.fwd_saved + ((int64_t)tile * kFwdSavedChunks + w.x_chunk0) * kAChunkBytes + qr * kDwPiece;
          const uint32_t dst = smem_base + s * kDwStageBytes;
#pragma unroll
          for (int cp = 0; cp < 8; ++cp) {   // cp = chunk*2 + plane; planes are 16 KiB apart inside a 32 KiB chunk
            bulk_g2s(dst + cp * kDwPiece, dy + (int64_t)cp * kPlaneBytes, kDwPiece, bar_full(s));
            bulk_g2s(dst + kDwOperand + cp * kDwPiece, x + (int64_t)cp * kPlaneBytes, kDwPiece, bar_full(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp waits, one elected lane issues (keeps the descriptors in uniform registers, see chain.cu)
    constexpr uint32_t idesc = umma_idesc_f16(128, 256, 1, 1);   // both operands MN-major
    constexpr uint32_t kLbo = 2 * kDwPiece;   // next 64-feature group of the same plane
    constexpr uint32_t kSbo = 1024;           // next 8 rays
    uint32_t git = 0;
    for (uint32_t k = 0;; ++k) {
      mbar_wait(bar(kDwBarItemFull + (k & 1u)), (k >> 1) & 1u);
      const int item = item_ring[k & 1u];
      if (item < 0) break;
      const int stages = decode(item).stages;
      if (k > 0) mbar_wait(bar(kDwBarTmemFree), (k - 1) & 1u);   // the epilogue has read the previous item's accumulators
      tc_fence_after_sync();
      for (int it = 0; it < stages; ++it, ++git) {
        const uint32_t s = git % kDwStages, ph = (git / kDwStages) & 1u;
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
              umma_f16(d, umma_desc_sw128(a_hi, kLbo, kSbo), umma_desc_sw128(b_hi, kLbo, kSbo), idesc, first);
              umma_f16(d, umma_desc_sw128(a_lo, kLbo, kSbo), umma_desc_sw128(b_hi, kLbo, kSbo), idesc, 1u);
              umma_f16(d, umma_desc_sw128(a_hi, kLbo, kSbo), umma_desc_sw128(b_lo, kLbo, kSbo), idesc, 1u);
            }
          }
          umma_commit(bar_empty(s));
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(bar(kDwBarDone));
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kDwBarItemEmpty + (k & 1u)));
    }
  } else if (warp >= 4) {
    const uint32_t t = (warp - 4) * 32 + lane;   // 0..127: owns output columns 2t, 2t+1 of dY for the bias sum
    const uint32_t c = t >> 5;                   // chunk of those columns
    const uint32_t kc = (2 * t) & 63;            // position inside the chunk
    const uint32_t q = warp & 3;
    const uint32_t tmem_row = tmem_base + ((q * 32u) << 16);
    constexpr int kUnitFloats = kWidth * kWidth + kWidth;            // dW [256][256] + db [256]
    uint32_t git = 0;
    for (uint32_t k = 0;; ++k) {
      mbar_wait(bar(kDwBarItemFull + (k & 1u)), (k >> 1) & 1u);
      const int item = item_ring[k & 1u];
      if (item < 0) break;
      const Item w = decode(item);
      float s0 = 0.f, s1 = 0.f;
      for (int it = 0; it < w.stages; ++it, ++git) {
        const uint32_t s = git % kDwStages, ph = (git / kDwStages) & 1u;
        mbar_wait(bar_full(s), ph);
        const uint8_t* dy = smem_gen + s * kDwStageBytes;
#pragma unroll 8
        for (uint32_t r = 0; r < kDwRowsPerStage; ++r) {
          const uint32_t off = r * 128u + ((((kc >> 3) ^ (r & 7u)) << 4) | ((kc & 7u) << 1));
          const uint32_t hi = *reinterpret_cast<const uint32_t*>(dy + (2 * c) * kDwPiece + off);
          const uint32_t lo = *reinterpret_cast<const uint32_t*>(dy + (2 * c + 1) * kDwPiece + off);
          const float2 fh = plane_word_to_float2(hi), fl = plane_word_to_float2(lo);
          s0 += fh.x + fl.x;
          s1 += fh.y + fl.y;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty(s));
      }
      // where the results go: straight into the gradient buffer, or (split unit) into this item's scratch slot
      // A piece of a split unit either adds its result into the (zeroed) gradient buffer with L2 reductions - the
      // default: no scratch traffic, no reduction pass, summation order of the <= 8 pieces not fixed - or, in the
      // deterministic mode, writes it to scratch for the unit's last piece to sum in index order.
      // the dY operands carry the backward's loss scale (chain.cu, backward prologue): divide it out, exactly
      const float unscale = p.bwd_scale ? __ldg(p.bwd_scale + 1) : 1.f;
      s0 *= unscale;
      s1 *= unscale;
      const bool atomic = w.splits > 1 && !p.deterministic;
      float* part = (w.splits > 1 && p.deterministic) ? p.partials + (int64_t)item * kUnitFloats : nullptr;
      auto head_feature = [&](int col) {
        const int chunk = 4 * w.hg + (col >> 6), slot = col & 63;
        return p.input_kind == kInputX ? (64 * chunk + slot < kInDim ? 64 * chunk + slot : -1) : fused_slot_to_feature(chunk, slot);
      };
      // bias gradient (head groups all compute the same sum; group 0 writes it)
      if (part) {
        part[kWidth * kWidth + 2 * t] = s0;
        part[kWidth * kWidth + 2 * t + 1] = s1;
      } else if (!w.is_head || w.hg == 0) {
        float* db = p.grads + (w.is_head ? kOffHeadB : off_body_b(w.layer)) + 2 * t;
        if (atomic) { red_add_f32(db, s0); red_add_f32(db + 1, s1); }
        else if (p.accumulate) { db[0] += s0; db[1] += s1; } else { db[0] = s0; db[1] = s1; }
      }
      // epilogue: both accumulators -> global
      mbar_wait(bar(kDwBarDone), k & 1u);
      tc_fence_after_sync();
      float* stage = reinterpret_cast<float*>(smem_gen + kDwSmemStage) + q * (32 * kDwStageRow);
      for (int half = 0; half < 2; ++half) {
        for (int c8 = 0; c8 < 8; ++c8) {
          uint32_t r[32];
          tmem_ld32(tmem_row + 256u * half + 32u * c8, r);
          tmem_ld_wait();
          if (half == 1 && c8 == 7) {
            // every accumulator column of this warp's lanes is in registers: hand TMEM back to the MMA warp
            tc_fence_before_sync();
            named_bar_sync(1, 128);
            if (t == 0) mbar_arrive(bar(kDwBarTmemFree));
          }
          // TMEM gives every lane one output row (32 consecutive columns); transposed through shared memory a warp
          // instruction covers whole 128-byte row segments instead of 32 rows x 16 bytes
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(stage + lane * kDwStageRow + 4 * i) =
                make_float4(__uint_as_float(r[4 * i]) * unscale, __uint_as_float(r[4 * i + 1]) * unscale,
                            __uint_as_float(r[4 * i + 2]) * unscale, __uint_as_float(r[4 * i + 3]) * unscale);
          __syncwarp();
          const int o0 = half * 128 + (int)q * 32;   // first output feature (row of dW) of this warp's block
          if (part || !w.is_head) {
            float* base = (part ? part : p.grads + off_body_w(w.layer)) + (int64_t)o0 * kWidth + 32 * c8 + 4 * (lane & 7);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int row = 4 * jj + (lane >> 3);
              float4 v = *reinterpret_cast<const float4*>(stage + row * kDwStageRow + 4 * (lane & 7));
              float* dst = base + (int64_t)row * kWidth;
              if (atomic) {
                red_add_v4_f32(dst, v.x, v.y, v.z, v.w);
              } else {
                if (!part && p.accumulate) { const float4 a = *reinterpret_cast<const float4*>(dst); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
                *reinterpret_cast<float4*>(dst) = v;
              }
            }
          } else {
            const int feat = head_feature(32 * c8 + lane);
            float* dst = p.grads + kOffHeadW + (int64_t)o0 * kInDim + feat;
            if (feat >= 0) {
#pragma unroll 8
              for (int j = 0; j < 32; ++j) {
                const float v = stage[j * kDwStageRow + lane];
                if (atomic) red_add_f32(dst + (int64_t)j * kInDim, v);
                else if (p.accumulate) dst[(int64_t)j * kInDim] += v;
                else dst[(int64_t)j * kInDim] = v;
              }
            }
          }
          __syncwarp();   // the staging block is rewritten by the next 32 columns
        }
      }
      if (part) {
        // ticket: the last piece of this unit to finish reduces all partials in index order (deterministic result);
        // 16 independent float4 loads per thread and piece in flight (the loop is L2-latency bound)
        __threadfence();   // this thread's partial sums are visible device-wide before the ticket is taken
        named_bar_sync(1, 128);
        if (t == 0) *ticket_smem = atomicAdd(p.tickets + w.ob, 1);
        named_bar_sync(1, 128);
        if (*ticket_smem == w.splits - 1) {
          __threadfence();
          const float* base = p.partials + (int64_t)p.unit_first[w.ob] * kUnitFloats;
          constexpr int kVecs = kUnitFloats / 4;   // 16448
          constexpr int kU = 8, kThreads = 128;
          for (int v0 = (int)t; v0 < kVecs; v0 += kU * kThreads) {
            float4 acc[kU];
#pragma unroll
            for (int j = 0; j < kU; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sp = 0; sp < w.splits; sp += 2) {
              float4 xa[kU], xb[kU];
              const bool two = sp + 1 < w.splits;
#pragma unroll
              for (int j = 0; j < kU; ++j) {
                const int v = v0 + kThreads * j;
                xa[j] = xb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v < kVecs) {
                  xa[j] = __ldcg(reinterpret_cast<const float4*>(base + (int64_t)sp * kUnitFloats) + v);
                  if (two) xb[j] = __ldcg(reinterpret_cast<const float4*>(base + (int64_t)(sp + 1) * kUnitFloats) + v);
                }
              }
#pragma unroll
              for (int j = 0; j < kU; ++j) {
                acc[j].x += xa[j].x; acc[j].y += xa[j].y; acc[j].z += xa[j].z; acc[j].w += xa[j].w;
                if (two) { acc[j].x += xb[j].x; acc[j].y += xb[j].y; acc[j].z += xb[j].z; acc[j].w += xb[j].w; }
              }
            }
#pragma unroll
            for (int j = 0; j < kU; ++j) {
              const int v = v0 + kThreads * j;
              if (v >= kVecs) continue;
              const int idx = 4 * v;
              const float a[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
              if (idx >= kWidth * kWidth) {
                if (!w.is_head || w.hg == 0) {
                  float* db = p.grads + (w.is_head ? kOffHeadB : off_body_b(w.layer)) + (idx - kWidth * kWidth);
#pragma unroll
                  for (int e = 0; e < 4; ++e) db[e] = p.accumulate ? db[e] + a[e] : a[e];
                }
              } else if (!w.is_head) {
                float4* wp = reinterpret_cast<float4*>(p.grads + off_body_w(w.layer) + idx);
                float4 o4 = acc[j];
                if (p.accumulate) { const float4 c4 = *wp; o4.x += c4.x; o4.y += c4.y; o4.z += c4.z; o4.w += c4.w; }
                *wp = o4;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int feat = head_feature((idx + e) & (kWidth - 1));
                  if (feat >= 0) {
                    float* wp = p.grads + kOffHeadW + (int64_t)((idx + e) >> 8) * kInDim + feat;
                    *wp = p.accumulate ? *wp + a[e] : a[e];
                  }
                }
              }
            }
          }
        }
        named_bar_sync(1, 128);   // ticket_smem is rewritten by the next item
      }
      if (p.times && t == 0) p.times[w.ob * 4 + 3] = global_timer_ns();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kDwBarItemEmpty + (k & 1u)));
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// tail.0.weight / tail.0.bias gradients: dW_t[c,j] = sum_n dlogit[n,c] (z_43 + h)[n,j]; 768+3 outputs, CUDA cores.
// A fixed number of blocks each sum their 128-ray tiles in order into registers and write one partial row; the last
// block to finish (ticket) adds the rows in block order and overwrites the gradient entries: no atomics on the data, the
// result is bit-reproducible.
constexpr int kTailBlocks = 256;
constexpr int kTailOutputs = kOutDim * kWidth + kOutDim;   // 771
__global__ void __launch_bounds__(256) r2l_tail_grad_kernel(const __grid_constant__ TailGradParams p) {
  __shared__ float dl[128][3];
  __shared__ int is_last;
  const int j = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, bsum = 0.f;
  const int num_tiles = (int)((p.n_rays + 127) / 128);
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t n0 = (int64_t)tile * 128;
    __syncthreads();   // the previous tile's dl is no longer read
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
    const int rows = (int)((p.n_rays - n0) < 128 ? (p.n_rays - n0) : 128);
    for (int r = 0; r < rows; ++r) {
      const float z = p.zf[(n0 + r) * kWidth + j];
      a0 = fmaf(dl[r][0], z, a0); a1 = fmaf(dl[r][1], z, a1); a2 = fmaf(dl[r][2], z, a2);
    }
    if (j < 3)
      for (int r = 0; r < rows; ++r) bsum += dl[r][j];
  }
  float* part = p.partials + (int64_t)blockIdx.x * kTailOutputs;
  part[j] = a0; part[kWidth + j] = a1; part[2 * kWidth + j] = a2;
  if (j < 3) part[3 * kWidth + j] = bsum;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(p.ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int o = j; o < kTailOutputs; o += 256) {
    float s = 0.f;
#pragma unroll 8
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(p.partials + (int64_t)b * kTailOutputs + o);
    p.grads[kOffTailW + o] = s;   // tail.0.weight [3,256] and tail.0.bias [3] are contiguous in the flat buffer
  }
  if (threadIdx.x == 0) *p.ticket = 0;
}

// Backward preamble (one block): zero the readiness flags / tickets / queue words of this call and choose the loss scale
// the backward chain runs on: S = 2^k with S * max |d rgb| in [2^9, 2^10).  With |d logit| <= |d rgb| / 4 and
// |W_tail| ~ 2^-4 the dY operands then peak near 2^4 and sit around 1: inside the range where the fp16 hi/lo planes carry
// 22 bits, with a factor of ~2^12 of head room to fp16's maximum for gradients that grow along the chain.
__global__ void __launch_bounds__(1024) r2l_bwd_prep_kernel(const float* __restrict__ grad_rgb, int64_t n, int* __restrict__ ready,
                                                            int ready_ints, float* __restrict__ scale_out) {
  __shared__ float warp_max[32];
  for (int i = threadIdx.x; i < ready_ints; i += blockDim.x) ready[i] = 0;
  float m = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(grad_rgb) + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  for (int64_t i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(__ldg(grad_rgb + i)));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = warp_max[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) {
      float s = 1.f;
      if (m > 0.f && m < 3.0e38f) {   // NaN / inf / all-zero gradients: no scaling
        int e;
        frexpf(m, &e);                // m = f * 2^e, f in [0.5, 1)
        int k = 10 - e;
        k = k > 100 ? 100 : (k < -100 ? -100 : k);
        s = ldexpf(1.f, k);
      }
      scale_out[0] = s;
      scale_out[1] = 1.f / s;
    }
  }
}

cudaError_t launch_bwd_prep(const float* grad_rgb, int64_t n_values, int* ready, int ready_ints, float* scale_out, cudaStream_t stream) {
  r2l_bwd_prep_kernel<<<1, 1024, 0, stream>>>(grad_rgb, n_values, ready, ready_ints, scale_out);
  return cudaGetLastError();
}

cudaError_t launch_dw(const DwParams& p, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(r2l_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDwSmemBytes);
  if (e != cudaSuccess) return e;
  r2l_dw_kernel<<<p.grid, kDwThreads, kDwSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_tail_grads(const TailGradParams& p, cudaStream_t stream) {
  const int tiles = (int)((p.n_rays + 127) / 128);
  r2l_tail_grad_kernel<<<tiles < kTailBlocks ? tiles : kTailBlocks, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace r2l
