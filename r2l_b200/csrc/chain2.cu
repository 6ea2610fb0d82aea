// CTA-PAIR variant of the chain kernel (see chain.cu for the algorithm and the warp roles).
//
// Two CTAs of a cluster (one SM pair) walk two neighbouring 128-ray tiles through the network together and issue every
// GEMM as ONE tcgen05.mma.cta_group::2 (M = 256: 128 rays per CTA, N = 256): each CTA stages only ITS half of every
// weight image (128 of the 256 output features), so
//   * the same 96 KiB of shared memory hold a 6-stage weight ring instead of 3 stages, and
//   * L2 -> SM weight traffic per SM halves (42 -> 21 B/cycle at full tensor rate).
// Differences to chain.cu: the leader CTA (cluster rank 0) issues the MMAs for both; its operand barriers count the
// epilogue warps of BOTH CTAs (the peer's warps arrive remotely, signal-only: their smem writes were already published to
// the async proxy by fence.proxy.async, so the leader waits with ordinary CTA-scope try_wait: a cluster-scope acquire
// costs ~180 cycles per wait even when the phase has long completed); the peer's warp 1 forwards "my weight half landed"; each CTA keeps local
// copies of the operand barriers for its own store warp in the training modes; tcgen05.commit multicasts to both CTAs.
#include "kernels.cuh"
#include "ptx.cuh"

namespace r2l {
namespace pair {

constexpr int kNumWStages = 6;                 // 6 x 16 KiB: this CTA's half (128 output features) of a weight image
constexpr int kWHalfBytes = kWImageBytes / 2;
constexpr int kEpiWarps = 16;
constexpr int kChainThreads = (4 + kEpiWarps) * 32;   // 640
constexpr uint32_t kSmemA = 0;
constexpr uint32_t kSmemW = kABytes;                                  // 131072
constexpr uint32_t kSmemBar = kSmemW + kNumWStages * kWHalfBytes;     // 229376
constexpr uint32_t kSmemTail = kSmemBar + 384;                        // 128 x 3 floats
constexpr uint32_t kSmemUsed = kSmemTail + kTileM * 3 * 4;            // 231168
constexpr uint32_t kChainSmemBytes = kSmemUsed + 1024;                // + alignment slack
static_assert(kChainSmemBytes <= 232448, "exceeds the 227 KiB dynamic shared memory limit");

constexpr uint32_t kTmemZ = 0;
constexpr uint32_t kTmemH = 256;

// barrier slots (8 bytes each) inside kSmemBar
enum : uint32_t {
  kBarWFull = 0,                          // [3] weight image landed            (TMA tx -> MMA)
  kBarWEmpty = kBarWFull + kNumWStages,   // [3] weight slot consumed           (MMA commit -> producer)
  kBarAFull = kBarWEmpty + kNumWStages,   // [4] A chunk written                (8 epilogue warps -> MMA, store warp)
  kBarAEmpty = kBarAFull + kAChunks,      // [4] A chunk consumed (head ring)    (MMA commit -> encoder)
  kBarASaved = kBarAEmpty + kAChunks,     // [4] A chunk copied out to HBM       (store warp -> epilogue)
  kBarAccFull = kBarASaved + kAChunks,    //     accumulator of a layer complete (MMA commit -> epilogue)
  kBarA0Sub = kBarAccFull + 1,            // [4] slot 0 is published per 16-column k-step (kBarAFull[0] is unused):
                                          //     the first GEMM instructions of a layer start after 1/16 of the epilogue
  kBarWPeer = kBarA0Sub + 4,              // [6] leader: the peer CTA's half of weight stage s landed        (peer warp 1)
  kBarAMma = kBarWPeer + kNumWStages,     // [8] leader: MMA-facing operand barriers, 32 arrivals = the 16 epilogue warps of BOTH
                                          //     CTAs (index ks: k-step ks of slot 0; index 4 + s: the whole chunk in slot s).
                                          //     kBarAFull / kBarA0Sub stay CTA-local (16 arrivals) for each CTA's store warp.
  kBarCount = kBarAMma + 8
};
static_assert(8 * kBarCount + 8 <= 384, "barrier block overflow");

// Quarter QT of a fused-order K chunk: 16 slots = 8 (sin, cos) pairs; the last quarter ends with x0,x1,x2,0.
template <int QT>
__device__ __forceinline__ void encode_quarter(const float (&x)[3], float (&out)[16]) {
  // fused feature order of layout.cuh: slot 2p = sin, 2p+1 = cos, pair p = coordinate*10 + frequency
#pragma unroll
  for (int i = 0; i < (QT < 3 ? 8 : 6); ++i) {
    const int p = QT * 8 + i;
    const int c = p / kFreqs, f = p % kFreqs;
    const float arg = __fmul_rn(x[c], static_cast<float>(1 << f));  // exact: power-of-two scale
    float s, co;
    sincosf(arg, &s, &co);
    out[2 * i] = s;
    out[2 * i + 1] = co;
  }
  if (QT == 3) {
    out[12] = x[0];
    out[13] = x[1];
    out[14] = x[2];
    out[15] = 0.f;
  }
}

// One 16-byte operand unit (8 consecutive K-values of one row), both planes.
__device__ __forceinline__ void store_a_unit(uint32_t a_chunk_addr, uint32_t row, uint32_t unit, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
  const uint32_t off = row * 128u + ((unit ^ (row & 7u)) << 4);
  st_shared_v4(a_chunk_addr + off, hi[0], hi[1], hi[2], hi[3]);
  st_shared_v4(a_chunk_addr + kPlaneBytes + off, lo[0], lo[1], lo[2], lo[3]);
}

template <int MODE>
__global__ void __launch_bounds__(kChainThreads, 1) r2l_chain_pair_kernel(const __grid_constant__ ChainParams p) {
  constexpr bool kIsBwd = MODE == kBwd;
  constexpr bool kSave = MODE != kFwdInfer;
  constexpr int kFirstChunks = kIsBwd ? kAChunks : kSamples;   // A chunks built before the first GEMM
  constexpr int kLayers = kIsBwd ? kBodyLayers : kBodyLayers + 1;  // GEMMs per tile
  constexpr int kImagesPerTile = kIsBwd ? 8 * kBodyLayers : 32 + 8 * kBodyLayers;
  // chunks saved per tile: forward 16 (PE) + 86*4 ; backward 4 (g_43) + 86*4
  constexpr int kSavedChunksPerTile = kFirstChunks + 4 * kBodyLayers;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + kSmemBar;
  auto bar = [&](uint32_t i) { return bar0 + 8u * i; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kSmemBar + 8 * kBarCount);
  float* tail_smem = reinterpret_cast<float*>(smem_gen + kSmemTail);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_ptiles = (p.num_tiles + 1) >> 1;    // tile pairs; CTA `rank` owns tile 2*pt + rank (may be a dummy)

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kNumWStages; ++i) {
      mbar_init(bar(kBarWFull + i), 1);
      mbar_init(bar(kBarWEmpty + i), 1);
    }
    for (int i = 0; i < kAChunks; ++i) {
      mbar_init(bar(kBarAFull + i), kEpiWarps);
      mbar_init(bar(kBarAEmpty + i), 1);
      mbar_init(bar(kBarASaved + i), 1);
      mbar_init(bar(kBarA0Sub + i), kEpiWarps);
    }
    mbar_init(bar(kBarAccFull), 1);
    for (int i = 0; i < kNumWStages; ++i) mbar_init(bar(kBarWPeer + i), 1);
    for (int i = 0; i < 8; ++i) mbar_init(bar(kBarAMma + i), 2 * kEpiWarps);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
    tmem_relinquish_pair();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();       // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ======================= weight producer =======================
    if (lane == 0) {
      const uint8_t* head_images =
          p.packed + (int64_t)(p.input_kind == kInputX ? kImgHeadNatural : kImgHeadFused) * kWImageBytes;
      const uint8_t* body_images = p.packed + (int64_t)(kIsBwd ? kImgBodyT : kImgBody) * kWImageBytes;
      uint32_t it = 0;
      long long t_wait = 0;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        const int tile = 2 * pt + (int)rank; (void)tile;
        for (int i = 0; i < kImagesPerTile; ++i, ++it) {
          const uint32_t ws = it % kNumWStages, ph = (it / kNumWStages) & 1u;
          const long long t0 = p.stats ? clock64() : 0;
          mbar_wait(bar(kBarWEmpty + ws), ph ^ 1u);
          if (p.stats) t_wait += clock64() - t0;
          mbar_arrive_expect_tx(bar(kBarWFull + ws), kWHalfBytes);
          const uint8_t* src;
          if (kIsBwd) src = body_images + (int64_t)i * kWImageBytes;
          else src = i < 32 ? head_images + (int64_t)i * kWImageBytes : body_images + (int64_t)(i - 32) * kWImageBytes;
          bulk_g2s(smem_base + kSmemW + ws * kWHalfBytes, src + rank * kWHalfBytes, kWHalfBytes, bar(kBarWFull + ws));
        }
      }
      if (p.stats) p.stats[blockIdx.x * 8 + 3] = t_wait;
    }
  } else if (warp == 1 && rank == 1) {
    // ======================= peer: forward "my weight half landed" to the leader =======================
    if (lane == 0) {
      const int per_tile = kIsBwd ? 8 * kBodyLayers : 32 + 8 * kBodyLayers;
      uint32_t it = 0;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        for (int i = 0; i < per_tile; ++i, ++it) {
          const uint32_t ws = it % kNumWStages;
          mbar_wait(bar(kBarWFull + ws), (it / kNumWStages) & 1u);
          mbar_arrive_cluster_relaxed(mapa_cluster(bar(kBarWPeer + ws), 0));
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA) =======================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256, 0, 0);   // M = 256 over the CTA pair
      uint32_t it = 0, a_phase = 0;
      long long t_a_head = 0, t_a_body = 0, t_w = 0;
      const long long t_begin = p.stats ? clock64() : 0;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        const int tile = 2 * pt + (int)rank; (void)tile;
        for (int l = 0; l < kLayers; ++l) {
          // forward: l = 0 head (16 chunks, -> Z fresh), odd l -> H fresh, even l -> Z accumulate
          // backward: j = l: even j (da = g W2) -> H fresh, odd j (g += dh W1) -> Z accumulate
          const int nkc = (!kIsBwd && l == 0) ? kSamples : kAChunks;
          const bool to_h = kIsBwd ? ((l & 1) == 0) : ((l & 1) != 0);
          const bool fresh = to_h || (!kIsBwd && l == 0);
          const bool ring = !kIsBwd && l == 0;   // head: A chunks recycle through the 4 slots
          const uint32_t d = tmem_base + (to_h ? kTmemH : kTmemZ);
          const bool tr = p.trace != nullptr && pt == pair_id;
          for (int kc = 0; kc < nkc; ++kc) {
            const uint32_t slot = kc & 3;
            const uint32_t a_hi = smem_base + kSmemA + slot * kAChunkBytes;
            const uint32_t a_lo = a_hi + kPlaneBytes;
            auto wait_w = [&](uint32_t i) {
              const long long t0 = p.stats ? clock64() : 0;
              mbar_wait(bar(kBarWFull + i % kNumWStages), (i / kNumWStages) & 1u);            // my half
              mbar_wait(bar(kBarWPeer + i % kNumWStages), (i / kNumWStages) & 1u);    // the peer's half
              if (p.stats) t_w += clock64() - t0;
              tc_fence_after_sync();
            };
            auto wait_a = [&](uint32_t barrier, uint32_t bit) {
              const long long t0 = p.stats ? clock64() : 0;
              (void)barrier;   // the 32-arrival MMA-facing barrier of this event
              mbar_wait(bar(kBarAMma + (bit >= 4 ? bit - 4 : 4 + bit)), (a_phase >> bit) & 1u);
              if (p.stats) { if (l == 0) t_a_head += clock64() - t0; else t_a_body += clock64() - t0; }
              a_phase ^= 1u << bit;
              tc_fence_after_sync();
            };
            if (slot == 0) {
              // k-step granular: the hi-image MMAs of a k-step are issued as soon as its 16 columns are published;
              // the W_hi slot is released after them, the lo-image MMAs follow (same issue order as the other chunks)
              wait_w(it);
              const uint32_t b_hi = smem_base + kSmemW + (it % kNumWStages) * kWHalfBytes;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                wait_a(kBarA0Sub + ks, 4 + ks);
                if (tr && kc == 0 && ks == 0) p.trace[((int64_t)blockIdx.x * 5 + 0) * 96 + l] = clock64();
                umma_bf16_pair(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b_hi + 32 * ks, 16, 1024), idesc,
                          (fresh && kc == 0 && ks == 0) ? 0u : 1u);
                umma_bf16_pair(d, umma_desc_sw128(a_lo + 32 * ks, 16, 1024), umma_desc_sw128(b_hi + 32 * ks, 16, 1024), idesc, 1u);
              }
              umma_commit_pair(bar(kBarWEmpty + it % kNumWStages));
              ++it;
              wait_w(it);
              const uint32_t b_lo = smem_base + kSmemW + (it % kNumWStages) * kWHalfBytes;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16_pair(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b_lo + 32 * ks, 16, 1024), idesc, 1u);
              umma_commit_pair(bar(kBarWEmpty + it % kNumWStages));
              ++it;
            } else {
              wait_a(kBarAFull + slot, slot);
              {  // W_hi image: A_hi*W_hi + A_lo*W_hi
                wait_w(it);
                const uint32_t b = smem_base + kSmemW + (it % kNumWStages) * kWHalfBytes;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_bf16_pair(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b + 32 * ks, 16, 1024), idesc, 1u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_bf16_pair(d, umma_desc_sw128(a_lo + 32 * ks, 16, 1024), umma_desc_sw128(b + 32 * ks, 16, 1024), idesc, 1u);
                umma_commit_pair(bar(kBarWEmpty + it % kNumWStages));
                ++it;
              }
              {  // W_lo image: A_hi*W_lo
                wait_w(it);
                const uint32_t b = smem_base + kSmemW + (it % kNumWStages) * kWHalfBytes;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_bf16_pair(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b + 32 * ks, 16, 1024), idesc, 1u);
                umma_commit_pair(bar(kBarWEmpty + it % kNumWStages));
                ++it;
              }
            }
            if (ring) umma_commit_pair(bar(kBarAEmpty + slot));
          }
          umma_commit_pair(bar(kBarAccFull));
          if (tr) p.trace[((int64_t)blockIdx.x * 5 + 1) * 96 + l] = clock64();
        }
        if constexpr (kIsBwd) {
          // the last backward epilogue publishes 4 more chunks (d head pre-activation, consumed only by the
          // store warp): step over those phases so the parity bookkeeping stays aligned for the next tile
          for (uint32_t ks = 0; ks < 4; ++ks) {
            mbar_wait(bar(kBarAMma + ks), (a_phase >> (4 + ks)) & 1u);
            a_phase ^= 1u << (4 + ks);
          }
          for (uint32_t slot = 1; slot < kAChunks; ++slot) {
            mbar_wait(bar(kBarAMma + 4 + slot), (a_phase >> slot) & 1u);
            a_phase ^= 1u << slot;
          }
        }
      }
      if (p.stats) {
        p.stats[blockIdx.x * 8 + 5] = global_timer_ns();
        p.stats[blockIdx.x * 8 + 0] = t_a_head;
        p.stats[blockIdx.x * 8 + 1] = t_a_body;
        p.stats[blockIdx.x * 8 + 2] = t_w;
        p.stats[blockIdx.x * 8 + 4] = clock64() - t_begin;
      }
    }
  } else if (warp == 3) {
    // ======================= operand-image store warp (train / bwd) =======================
    if (kSave && lane == 0) {
      // Up to four 32 KiB stores in flight: chunk i's slot is released (kBarASaved) once store i has finished
      // READING shared memory, which we learn when at most 3 younger bulk groups are still pending.  The
      // epilogue rewrites a slot exactly 4 chunks after it published it, so this lag can never deadlock.
      uint32_t a_phase = 0;
      int64_t issued = 0;
      const bool signal = kIsBwd && p.ready != nullptr;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        const int tile = 2 * pt + (int)rank; (void)tile;
        uint8_t* dst = p.saved + (int64_t)tile * kSavedChunksPerTile * kAChunkBytes;
        int signalled = 0;   // operand groups (4 chunks = one layer's dY) of this tile already announced
        for (int i = 0; i < kSavedChunksPerTile; ++i, ++issued) {
          const uint32_t slot = i & 3;   // first chunks cycle the ring; body chunk c lives in slot c
          if (slot == 0) {
            // k-steps 0/2 and 1/3 of slot 0 are written by different warps: the chunk is complete when the last
            // k-step of both groups has been published
            mbar_wait(bar(kBarA0Sub + 2), a_phase & 1u);
            mbar_wait(bar(kBarA0Sub + 3), a_phase & 1u);
          } else {
            mbar_wait(bar(kBarAFull + slot), (a_phase >> slot) & 1u);
          }
          a_phase ^= 1u << slot;
          bulk_s2g(dst + (int64_t)i * kAChunkBytes, smem_base + kSmemA + slot * kAChunkBytes, kAChunkBytes);
          bulk_commit();
          if (issued >= 3) {
            bulk_wait_read<3>();         // store (issued - 3) has read its slot
            mbar_arrive(bar(kBarASaved + ((slot + 1) & 3)));
          }
          if (signal && tile < p.num_tiles && slot == 3) {
            // everything but the 4 stores just issued has landed in global memory: announce those groups so the
            // weight-gradient kernel (running concurrently on idle SMs) may start on their layers
            bulk_wait_all<4>();
            for (; signalled < (i >> 2); ++signalled) flag_release_add(p.ready + signalled);
          }
        }
        if (signal && tile < p.num_tiles) {   // a dummy tile (odd tile count) must not be counted
          bulk_wait_all<0>();
          for (; signalled < kSavedChunksPerTile / 4; ++signalled) flag_release_add(p.ready + signalled);
        }
      }
      // drain: release the last three slots in issue order, then wait for the writes to land
      const int64_t tail = issued < 3 ? issued : 3;
      bulk_wait_read<0>();
      for (int64_t k = 0; k < tail; ++k) mbar_arrive(bar(kBarASaved + (uint32_t)((issued - tail + k) & 3)));
      bulk_wait_all<0>();
    }
  } else if (warp >= 4) {
    // ======================= encoder / epilogue =======================
    const uint32_t ew = warp - 4;
    const uint32_t q = ew & 3u;       // TMEM lane quarter (== warp % 4)
    const uint32_t qt = ew >> 2;      // which quarter of a row's columns this thread works on (0..3)
    const uint32_t row = q * 32u + lane;
    const uint32_t tmem_row = tmem_base + ((q * 32u) << 16);
    const float* cumbias = reinterpret_cast<const float*>(p.packed + kPackOffCumBias);
    const float* headb = reinterpret_cast<const float*>(p.packed + kPackOffHeadB);
    const float* b1 = reinterpret_cast<const float*>(p.packed + kPackOffB1);
    const float* tailw = reinterpret_cast<const float*>(p.packed + kPackOffTailW);
    const float* tailb = reinterpret_cast<const float*>(p.packed + kPackOffTailB);
    float* hrow = p.scratch + ((int64_t)blockIdx.x * kTileM + row) * kWidth;
    uint32_t acc_phase = 0;
    uint32_t saved_phase = 0;   // per-slot parity of kBarASaved
    (void)cumbias; (void)headb; (void)b1; (void)tailb; (void)tail_smem;
    // Epilogue column ownership inside a 64-column chunk: k-steps g0 = qt>>1 and g0+2, and inside each k-step the
    // 8-column unit u = qt&1, i.e. the 16-byte operand units 2g+u.  K-steps 0/2 belong to the warps with qt in {0,1},
    // k-steps 1/3 to qt in {2,3}: the first 16 columns of a layer's output are ready after 8 warps did 8 columns each.
    const uint32_t g0 = qt >> 1, uu = qt & 1u;

    // kBarA0Sub[ks] counts all 16 warps: owners arrive when their part of k-step ks is written, the others at once.
    auto arrive_mma = [&](uint32_t idx) {   // leader: local arrive; peer: signal-only remote arrive on the leader's barrier
      if (lane == 0) {
        if (rank == 0) mbar_arrive(bar(kBarAMma + idx));
        else mbar_arrive_cluster_relaxed(mapa_cluster(bar(kBarAMma + idx), 0));
      }
    };
    auto arrive_sub = [&](uint32_t ks) {
      arrive_mma(ks);
      if (kSave && lane == 0) mbar_arrive(bar(kBarA0Sub + ks));
    };
    auto make_visible = [&]() {   // generic-proxy smem writes -> tensor core (async proxy), TMEM reads ordered
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
    };
    auto publish = [&](uint32_t slot) {          // this warp's part of a whole 64-column chunk is written
      make_visible();
      if (slot == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) arrive_sub(ks);
      } else {
        arrive_mma(4 + slot);
        if (kSave && lane == 0) mbar_arrive(bar(kBarAFull + slot));
      }
    };
    // before rewriting a slot in save modes: the store warp must have copied the previous content out
    auto wait_saved = [&](uint32_t slot, bool first_use) {
      if (kSave) {
        if (!first_use) mbar_wait(bar(kBarASaved + slot), (saved_phase >> slot) & 1u);
        if (!first_use) saved_phase ^= 1u << slot;
      }
    };

    bool first_tile = true;
    for (int pt = pair_id; pt < num_ptiles; pt += num_pairs, first_tile = false) {
      const int tile = 2 * pt + (int)rank;
      const int64_t grow = (int64_t)tile * kTileM + row;
      const bool valid = tile < p.num_tiles && grow < p.n_rays;

      if constexpr (!kIsBwd) {
        // ---- head A operand: 16 chunks through the 4-slot A ring; thread (row, qt) writes slots 16 qt .. 16 qt + 15 ----
        float o[3] = {0.f, 0.f, 0.f}, dd[3] = {0.f, 0.f, 0.f};
        if (p.input_kind == kInputRays && valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            o[c] = __ldg(p.in0 + grow * 3 + c);
            dd[c] = __ldg(p.in1 + grow * 3 + c);
          }
        }
        for (int c = 0; c < kSamples; ++c) {
          const uint32_t slot = c & 3;
          float f[16];
          if (p.input_kind == kInputX) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int feat = 64 * c + 16 * (int)qt + i;
              f[i] = (valid && feat < kInDim) ? __ldg(p.in0 + grow * kInDim + feat) : 0.f;
            }
          } else {
            float x[3] = {0.f, 0.f, 0.f};
            if (valid) {
              if (p.input_kind == kInputPts) {
#pragma unroll
                for (int k = 0; k < 3; ++k) x[k] = __ldg(p.in0 + grow * (3 * kSamples) + 3 * c + k);
              } else {
                float z = p.z_lo[c];
                if (p.t_rand != nullptr)  // lower + (upper - lower) * t_rand, nerf_raybased.py:123
                  z = __fadd_rn(z, __fmul_rn(p.z_diff[c], __ldg(p.t_rand + grow * kSamples + c)));
#pragma unroll
                for (int k = 0; k < 3; ++k) x[k] = __fadd_rn(o[k], __fmul_rn(dd[k], z));  // :124
              }
            }
            if (qt == 0) encode_quarter<0>(x, f);
            else if (qt == 1) encode_quarter<1>(x, f);
            else if (qt == 2) encode_quarter<2>(x, f);
            else encode_quarter<3>(x, f);
          }
          if (c >= 4) mbar_wait(bar(kBarAEmpty + slot), ((c >> 2) - 1) & 1u);
          wait_saved(slot, first_tile && c < 4);
          const uint32_t chunk_addr = smem_base + kSmemA + slot * kAChunkBytes;
          store_a_unit(chunk_addr, row, 2 * qt, &f[0]);
          store_a_unit(chunk_addr, row, 2 * qt + 1, &f[8]);
          publish(slot);
        }
      } else {
        // ---- backward prologue: d logit = d rgb * rgb (1 - rgb);  g = d logit . W_tail  (dL/dz_43 = dL/dh_skip) ----
        float dl[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float y = __ldg(p.rgb_in + grow * 3 + c);
            dl[c] = __ldg(p.grad_rgb + grow * 3 + c) * y * (1.f - y);
          }
        }
        for (int c = 0; c < kAChunks; ++c) {
          const uint32_t col = 64u * c + 16u * qt;   // 16 consecutive columns per thread here
          float v[16];
          uint32_t r[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(tailw + col) + i);
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(tailw + kWidth + col) + i);
            const float4 w2 = __ldg(reinterpret_cast<const float4*>(tailw + 2 * kWidth + col) + i);
            v[4 * i + 0] = fmaf(dl[2], w2.x, fmaf(dl[1], w1.x, dl[0] * w0.x));
            v[4 * i + 1] = fmaf(dl[2], w2.y, fmaf(dl[1], w1.y, dl[0] * w0.y));
            v[4 * i + 2] = fmaf(dl[2], w2.z, fmaf(dl[1], w1.z, dl[0] * w0.z));
            v[4 * i + 3] = fmaf(dl[2], w2.w, fmaf(dl[1], w1.w, dl[0] * w0.w));
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
          tmem_st8(tmem_row + kTmemZ + col, &r[0]);
          tmem_st8(tmem_row + kTmemZ + col + 8, &r[8]);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(hrow + col)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          tmem_st_wait();
          wait_saved(c, first_tile);
          const uint32_t chunk_addr = smem_base + kSmemA + c * kAChunkBytes;
          store_a_unit(chunk_addr, row, 2 * qt, &v[0]);
          store_a_unit(chunk_addr, row, 2 * qt + 1, &v[8]);
          publish(c);
        }
      }

      // ---- layer epilogues ----
      float dot0 = 0.f, dot1 = 0.f, dot2 = 0.f;
      for (int l = 0; l < kLayers; ++l) {
        const bool from_h = kIsBwd ? ((l & 1) == 0) : ((l & 1) != 0);
        const bool last = l == kLayers - 1;
        // forward tables
        const bool relu = !kIsBwd && ((l == 0) || from_h);
        const float* bias = nullptr;
        if constexpr (!kIsBwd)
          bias = l == 0 ? headb : ((l & 1) ? b1 + (l >> 1) * kWidth : cumbias + (l >> 1) * kWidth);
        // backward: mask source = hi plane of a saved forward operand image.
        //   even j (da -> dh): a_k with k = 42 - j/2, forward saved chunk index 16 + 4*(2k+1) + c
        //   last (j = 85):     h (= A_z(0), forward saved chunks 16 + c), applied to g_0 + dL/dz_43
        const uint8_t* mask_img = nullptr;
        const bool masked = kIsBwd && (from_h || last);
        if constexpr (kIsBwd) {
          const int k = (kBlocks - 1) - (l >> 1);
          const int64_t chunk0 = from_h ? (kSamples + 4 * (2 * k + 1)) : kSamples;
          mask_img = p.fwd_saved + ((int64_t)tile * kFwdSavedChunks + chunk0) * kAChunkBytes;
        }
        // side data of a chunk (bias / ReLU mask of my two 8-column units); chunk 0's is fetched while the MMAs run
        float4 bq[4];
        uint4 mq[2];
        auto load_side = [&](int c) {
          if constexpr (!kIsBwd) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4* b4 = reinterpret_cast<const float4*>(bias + 64 * c + 16 * (g0 + 2 * h) + 8 * uu);
              bq[2 * h] = __ldg(b4);
              bq[2 * h + 1] = __ldg(b4 + 1);
            }
          } else {
            if (masked) {
              const uint8_t* plane = mask_img + (int64_t)c * kAChunkBytes;
#pragma unroll
              for (int h = 0; h < 2; ++h)
                mq[h] = __ldg(reinterpret_cast<const uint4*>(plane + row * 128u + (((2u * (g0 + 2 * h) + uu) ^ (row & 7u)) << 4)));
            }
          }
        };
        load_side(0);
        mbar_wait(bar(kBarAccFull), acc_phase);
        acc_phase ^= 1u;
        tc_fence_after_sync();
        const bool tr = p.trace != nullptr && pt == pair_id && warp == 4 && lane == 0;
        if (tr) p.trace[((int64_t)blockIdx.x * 5 + 2) * 96 + l] = clock64();
        const bool feeds_mma = !last;                     // the last epilogue of a tile produces no further GEMM input
        const bool produces_chunk = feeds_mma || kIsBwd;  // backward's last output (d head pre-activation) is saved for dw
        for (int c = 0; c < kAChunks; ++c) {
          uint32_t r[16];
          const uint32_t tacc = tmem_row + (from_h ? kTmemH : kTmemZ) + 64u * c + 8u * uu;
          tmem_ld8(tacc + 16u * g0, &r[0]);
          tmem_ld8(tacc + 16u * (g0 + 2), &r[8]);
          if (c > 0) load_side(c);
          if (produces_chunk) {
            wait_saved(c, false);
            if (c == 0) { arrive_sub(1 - g0); arrive_sub(3 - g0); }   // the k-steps of chunk 0 I do not write
          }
          tmem_ld_wait();
          const uint32_t chunk_addr = smem_base + kSmemA + c * kAChunkBytes;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t g = g0 + 2 * h;
            const uint32_t col = 64u * c + 16u * g + 8u * uu;
            float v[8];
            if constexpr (!kIsBwd) {
              v[0] = __uint_as_float(r[8 * h + 0]) + bq[2 * h].x;
              v[1] = __uint_as_float(r[8 * h + 1]) + bq[2 * h].y;
              v[2] = __uint_as_float(r[8 * h + 2]) + bq[2 * h].z;
              v[3] = __uint_as_float(r[8 * h + 3]) + bq[2 * h].w;
              v[4] = __uint_as_float(r[8 * h + 4]) + bq[2 * h + 1].x;
              v[5] = __uint_as_float(r[8 * h + 5]) + bq[2 * h + 1].y;
              v[6] = __uint_as_float(r[8 * h + 6]) + bq[2 * h + 1].z;
              v[7] = __uint_as_float(r[8 * h + 7]) + bq[2 * h + 1].w;
              if (relu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
              }
              if (l == 0) {
                // z_0 = h: seed the TMEM residual stream and keep h for the outer skip (:543)
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(v[i]);
                tmem_st8(tmem_row + kTmemZ + col, w);
                reinterpret_cast<float4*>(hrow + col)[0] = make_float4(v[0], v[1], v[2], v[3]);
                reinterpret_cast<float4*>(hrow + col)[1] = make_float4(v[4], v[5], v[6], v[7]);
                tmem_st_wait();
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * h + i]);
              if (last) {  // + dL/dz_43 through the outer skip
                const float4 s0 = reinterpret_cast<const float4*>(hrow + col)[0];
                const float4 s1 = reinterpret_cast<const float4*>(hrow + col)[1];
                v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
                v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
              }
              if (masked) {
                // ReLU mask from the hi plane of the saved forward operand (a > 0  <=>  bf16 hi != 0; a >= 0 always)
                const uint32_t w[4] = {mq[h].x, mq[h].y, mq[h].z, mq[h].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if ((w[e] & 0x00007FFFu) == 0u) v[2 * e] = 0.f;
                  if ((w[e] & 0x7FFF0000u) == 0u) v[2 * e + 1] = 0.f;
                }
              }
            }
            if (produces_chunk) {
              store_a_unit(chunk_addr, row, 2 * g + uu, v);
              if (c == 0) {
                make_visible();
                arrive_sub(g);
                if (tr && g == 0) p.trace[((int64_t)blockIdx.x * 5 + 3) * 96 + l] = clock64();
              }
            }
            if constexpr (!kIsBwd) {
              if (last) {
                // tail: rgb = sigmoid(W_t (z_43 + h) + b_t), partial dot over my 8 columns
                float zf[8];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const float4 h4 = reinterpret_cast<const float4*>(hrow + col)[i];
                  const float4 w0 = __ldg(reinterpret_cast<const float4*>(tailw + col) + i);
                  const float4 w1 = __ldg(reinterpret_cast<const float4*>(tailw + kWidth + col) + i);
                  const float4 w2 = __ldg(reinterpret_cast<const float4*>(tailw + 2 * kWidth + col) + i);
                  const float z0 = v[4 * i] + h4.x, z1 = v[4 * i + 1] + h4.y, z2 = v[4 * i + 2] + h4.z, z3 = v[4 * i + 3] + h4.w;
                  zf[4 * i] = z0; zf[4 * i + 1] = z1; zf[4 * i + 2] = z2; zf[4 * i + 3] = z3;
                  dot0 = fmaf(z0, w0.x, dot0); dot0 = fmaf(z1, w0.y, dot0); dot0 = fmaf(z2, w0.z, dot0); dot0 = fmaf(z3, w0.w, dot0);
                  dot1 = fmaf(z0, w1.x, dot1); dot1 = fmaf(z1, w1.y, dot1); dot1 = fmaf(z2, w1.z, dot1); dot1 = fmaf(z3, w1.w, dot1);
                  dot2 = fmaf(z0, w2.x, dot2); dot2 = fmaf(z1, w2.y, dot2); dot2 = fmaf(z2, w2.z, dot2); dot2 = fmaf(z3, w2.w, dot2);
                }
                if (MODE == kFwdTrain && valid) {  // z_43 + h for the tail weight gradient
                  float* zrow = p.zf_out + grow * kWidth + col;
                  reinterpret_cast<float4*>(zrow)[0] = make_float4(zf[0], zf[1], zf[2], zf[3]);
                  reinterpret_cast<float4*>(zrow)[1] = make_float4(zf[4], zf[5], zf[6], zf[7]);
                }
              }
            }
          }
          if (produces_chunk && c > 0) publish(c);
        }
        if (tr) p.trace[((int64_t)blockIdx.x * 5 + 4) * 96 + l] = clock64();
      }
      if constexpr (!kIsBwd) {
        // combine the four column-quarter partial sums of each ray through TMEM (the H region is idle here, and the
        // four threads of a ray share its TMEM lane): fixed summation order, no shared memory needed
        uint32_t w4[4] = {__float_as_uint(dot0), __float_as_uint(dot1), __float_as_uint(dot2), 0u};
        tmem_st4(tmem_row + kTmemH + 4u * qt, w4);
        tmem_st_wait();
        tc_fence_before_sync();
        named_bar_sync(1, kEpiWarps * 32);
        tc_fence_after_sync();
        if (qt == 0) {
          uint32_t a[16];
          tmem_ld16(tmem_row + kTmemH, a);
          tmem_ld_wait();
          if (valid) {
            const float s0 = ((__uint_as_float(a[0]) + __uint_as_float(a[4])) + __uint_as_float(a[8])) + __uint_as_float(a[12]) + __ldg(tailb + 0);
            const float s1 = ((__uint_as_float(a[1]) + __uint_as_float(a[5])) + __uint_as_float(a[9])) + __uint_as_float(a[13]) + __ldg(tailb + 1);
            const float s2 = ((__uint_as_float(a[2]) + __uint_as_float(a[6])) + __uint_as_float(a[10])) + __uint_as_float(a[14]) + __ldg(tailb + 2);
            p.rgb[grow * 3 + 0] = 1.f / (1.f + expf(-s0));
            p.rgb[grow * 3 + 1] = 1.f / (1.f + expf(-s1));
            p.rgb[grow * 3 + 2] = 1.f / (1.f + expf(-s2));
          }
        }
        tc_fence_before_sync();
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();       // the leader's MMAs touch the peer's smem / TMEM until the very end
  tc_fence_after_sync();
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace pair
using namespace pair;

template <int MODE>
static cudaError_t launch_chain_pair_mode(const ChainParams& p, int grid, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(r2l_chain_pair_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kChainSmemBytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);          // even: CTAs 2i and 2i+1 form a pair
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = kChainSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, r2l_chain_pair_kernel<MODE>, p);
}

cudaError_t launch_chain_pair(int mode, const ChainParams& p, int grid, cudaStream_t stream) {
  switch (mode) {
    case kFwdInfer: return launch_chain_pair_mode<kFwdInfer>(p, grid, stream);
    case kFwdTrain: return launch_chain_pair_mode<kFwdTrain>(p, grid, stream);
    case kBwd: return launch_chain_pair_mode<kBwd>(p, grid, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace r2l
