// The "chain" kernel: one persistent CTA per SM walks a 128-ray tile through the whole 88-layer network
// without the activations ever leaving the SM.  Three instantiations share the machinery:
//
//   kFwdInfer  rays -> 16 points -> sin/cos encoding -> head -> 43 ResMLP blocks -> +h -> tail -> sigmoid
//   kFwdTrain  same, and every A operand it builds (the input of every Linear) is bulk-stored to HBM as a
//              ready-made tensor-core operand image for the weight-gradient kernel (dw.cu)
//   kBwd       d rgb -> d logit -> g = dL/dz_43 -> for k = 42..0 : da = g W2_k ; dh = da * (a_k > 0) ;
//              g += dh W1_k  -> (g + dL/dz_43) * (h > 0); every dY operand is stored for dw.cu
//
// Reference semantics: /root/reference/model/nerf_raybased.py — PointSampler.sample_train :114-126,
// PositionalEmbedder.__call__ :198-208, NeRF_v3_2.forward :539-544, ResMLP.forward :461-465; the backward
// is what torch.autograd derives for that graph (loss.backward(), main.py:1404).
//
// Per CTA (640 threads):
//   warp 0      weight producer: streams fp16 operand images (TMA bulk copy) into a 96 KiB ring
//   warp 1      MMA issuer: tcgen05.mma N=256,K=16, fp16x3 split (hi*hi + lo*hi + hi*lo), fp32 in TMEM
//   warp 2      TMEM allocator (512 columns: Z [0,256) = residual stream / running gradient,
//               H [256,512) = block hidden / its gradient)
//   warp 3      (train / bwd) operand-image store warp: smem A chunk -> HBM via bulk async stores
//   warps 4-19  epilogue / encoder: FOUR threads per ray (16 warps = 4 per SM sub-partition, so that TMEM-read and
//               fence latencies of one warp hide behind the arithmetic of the others). They build the first A
//               operand (positional encoding, or dL/dz_43) and after every layer turn the fp32 accumulator into the
//               next layer's fp16 hi/lo A operand in shared memory (bias/ReLU or mask, split, 128B swizzle).
//
// Three launch forms of the same kernel (template parameter FORM):
//   single  one CTA per 128-ray tile; 3 x 32 KiB weight ring; tcgen05.mma.cta_group::1, M = 128.
//   half    a cluster of two CTAs (one SM pair) shares ONE 128-ray tile: 64 rays per CTA, every GEMM is one
//           tcgen05.mma.cta_group::2 with M = 128 (64 rows from each CTA), which the tensor cores of the two SMs execute in
//           64 cycles instead of 128.  Each CTA's accumulator then lives in the "2x2" TMEM layout: row m, feature j at
//           lane m + 64 * (j / 128), column j % 128 - so the 16 epilogue warps give EIGHT threads to every ray (warps
//           of sub-partitions 0,1 own features 0..127 = K chunks 0,1 of the next layer, sub-partitions 2,3 own chunks
//           2,3) and the per-layer epilogue, which bounds the single form, halves.  Weight halves as in the pair form
//           (8 x 16 KiB ring).  The saved operand images keep the 128-row layout of the other forms: CTA r stores rows
//           64 r .. 64 r + 63 of each plane.  This is the form for small batches (the 4096-ray training step uses
//           64 SMs instead of 32 and each layer takes about half the time).
//   pair    a cluster of two CTAs (one SM pair) walks two neighbouring tiles together and issues every GEMM as ONE
//           tcgen05.mma.cta_group::2 (M = 256: 128 rays per CTA): each CTA stages only ITS half (128 output features)
//           of every weight image - 6 x 16 KiB ring, half the L2 -> SM weight traffic.  The leader CTA (cluster rank 0)
//           issues the MMAs for both; its MMA-facing operand barriers count the epilogue warps of BOTH CTAs (the peer's
//           warps arrive remotely, signal-only: their smem writes were already published to the async proxy by
//           fence.proxy.async, and a cluster-scope release would cost each warp hundreds of cycles); the peer's warp 1
//           forwards "my weight half landed"; tcgen05.commit multicasts to both CTAs.  Bit-identical results.
//           Measured: +3.5 % on the training step, -10 % on plain inference (c_api.cu picks the form per mode).
//
// TMEM residency of the residual stream: Z holds z_k - sum_{j<k} b2_j (forward) or g_k (backward) in fp32;
// the second GEMM of each block accumulates straight onto it, so the skip connection costs nothing.
// The head accumulates into Z as well (its epilogue rewrites Z in place with relu(h)): H is overwritten by
// block 0's first Linear while the head epilogue is still draining.
#include "kernels.cuh"
#include "ptx.cuh"

namespace r2l {

constexpr int kMaxWStages = 8;   // barrier slots reserved; single: 3 x 32 KiB, pair: 6 x 16 KiB, half: 8 x 16 KiB
constexpr int kEpiWarps = 16;
constexpr int kChainThreads = (4 + kEpiWarps) * 32;   // 640
// Shared-memory geometry of a form: [A ring: 4 slots x (hi plane | lo plane)] [weight ring] [barriers] [tail partials]
struct ChainGeom {
  uint32_t rows, plane, slot, w_stages, w_stage_bytes, off_w, off_bar, off_tail, used, tmem_h;
};
constexpr uint32_t kBarBlockBytes = 512;
__host__ __device__ constexpr ChainGeom chain_geom(int form) {
  ChainGeom g{};
  g.rows = form == kFormHalf ? 64u : 128u;              // rays per CTA
  g.plane = g.rows * 128u;                              // one fp16 plane of a 64-feature K chunk (SW128 K-major rows)
  g.slot = 2u * g.plane;
  g.w_stages = form == kFormHalf ? 8u : form == kFormPair ? 6u : 3u;
  g.w_stage_bytes = form == kFormSingle ? (uint32_t)kWImageBytes : (uint32_t)kWImageBytes / 2u;
  g.off_w = 4u * g.slot;
  g.off_bar = g.off_w + g.w_stages * g.w_stage_bytes;
  g.off_tail = g.off_bar + kBarBlockBytes;
  g.used = g.off_tail + (form == kFormHalf ? 64u * 16u * 3u * 4u : 0u);  // half: 16 partial tail dots per ray
  g.tmem_h = form == kFormHalf ? 128u : 256u;           // half: an accumulator is 128 columns wide (2x2 layout)
  return g;
}
__host__ __device__ constexpr uint32_t chain_smem_bytes(int form) { return chain_geom(form).used + 1024u; }   // + alignment slack
static_assert(chain_smem_bytes(kFormSingle) <= 232448 && chain_smem_bytes(kFormPair) <= 232448 &&
              chain_smem_bytes(kFormHalf) <= 232448, "exceeds the 227 KiB dynamic shared memory limit");

constexpr uint32_t kTmemZ = 0;
// half form only (its accumulators are 128 columns wide, so TMEM has room): Y = [256, 384), the accumulator of every GEMM
// whose result joins the residual stream (the head, the second Linear of a block, g += dh W1 in the backward).  The tensor
// core truncates its fp32 accumulator after every MMA (round toward zero; measured: DESIGN.md section 4 "Precision"), so
// accumulating 43 blocks x 48 MMAs straight onto the stream biases it by ~1e-4.  In the half form every GEMM therefore starts
// from zero, its small split terms are issued first, and the epilogue adds the result to the stream in fp32 (round to
// nearest); Z is then plain storage for the stream (tcgen05.ld / st by the epilogue threads).
constexpr uint32_t kTmemY = 256;

// barrier slots (8 bytes each) inside kSmemBar
enum : uint32_t {
  kBarWFull = 0,                          // [3] weight image landed            (TMA tx -> MMA)
  kBarWEmpty = kBarWFull + kMaxWStages,   // [3] weight slot consumed           (MMA commit -> producer)
  kBarAFull = kBarWEmpty + kMaxWStages,   // [4] A chunk written                (8 epilogue warps -> MMA, store warp)
  kBarAEmpty = kBarAFull + kAChunks,      // [4] A chunk consumed (head ring)    (MMA commit -> encoder)
  kBarASaved = kBarAEmpty + kAChunks,     // [4] A chunk copied out to HBM       (store warp -> epilogue)
  kBarAccFull = kBarASaved + kAChunks,    //     accumulator of a layer complete (MMA commit -> epilogue)
  kBarA0Sub = kBarAccFull + 1,            // [4] slot 0 is published per 16-column k-step (kBarAFull[0] is unused):
                                          //     the first GEMM instructions of a layer start after 1/16 of the epilogue
  kBarWPeer = kBarA0Sub + 4,              // [8] pair forms, leader: the peer CTA's half of weight stage s landed (peer warp 1)
  kBarAMma = kBarWPeer + kMaxWStages,     // [16] pair forms, leader: MMA-facing operand barriers, arrivals from the epilogue warps of
                                          //     BOTH CTAs; kBarAFull / kBarA0Sub then stay CTA-local for each CTA's store warp.
                                          //     pair: 8 used, 32 arrivals (index ks: k-step ks of slot 0; 4 + s: whole chunk in slot s)
                                          //     half: 16 used, 16 arrivals (index 4 s + ks: k-step ks of slot s; a slot is written
                                          //           by the 8 warps per CTA that own its feature half)
  kBarCount = kBarAMma + 16
};
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
__device__ __forceinline__ void store_a_unit(uint32_t a_chunk_addr, uint32_t row, uint32_t unit, const float* v,
                                             uint32_t plane_bytes = kPlaneBytes) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
  const uint32_t off = row * 128u + ((unit ^ (row & 7u)) << 4);
  st_shared_v4(a_chunk_addr + off, hi[0], hi[1], hi[2], hi[3]);
  st_shared_v4(a_chunk_addr + plane_bytes + off, lo[0], lo[1], lo[2], lo[3]);
}

template <int MODE, int FORM>
__global__ void __launch_bounds__(kChainThreads, 1) r2l_chain_kernel(const __grid_constant__ ChainParams p) {
  constexpr bool kIsBwd = MODE == kBwd;
  constexpr bool PAIR = FORM != kFormSingle;   // a cluster of two CTAs, the leader issues cta_group::2 MMAs for both
  constexpr bool HALF = FORM == kFormHalf;     // ... which share one 128-ray tile (64 rays each) instead of owning one each
  constexpr ChainGeom G = chain_geom(FORM);
  constexpr int kNumWStages = (int)G.w_stages;
  constexpr uint32_t kWStageBytes = G.w_stage_bytes;   // pair forms: my half of the output features
  constexpr uint32_t kSmemA = 0, kSmemW = G.off_w, kSmemBar = G.off_bar, kSmemTail = G.off_tail;
  constexpr uint32_t kSlotBytes = G.slot, kPlane = G.plane, kTmemH = G.tmem_h;
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
  // pair form: rank 0 = leader (issues the MMAs), 1 = peer; CTA `rank` of pair `pair_id` owns tile 2 * pt + rank (possibly a
  // dummy one past the end when the tile count is odd).  half form: both CTAs work on tile pt, CTA `rank` on its rows
  // 64 * rank .. 64 * rank + 63.  Single form: rank 0, one tile per step.
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int pair_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_pairs = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_ptiles = (PAIR && !HALF) ? (p.num_tiles + 1) >> 1 : p.num_tiles;
  auto tile_of = [&](int pt) { return (PAIR && !HALF) ? 2 * pt + (int)rank : pt; };   // 128-ray tile (= saved-image tile)
  const uint32_t row0 = HALF ? 64u * rank : 0u;   // my first row inside that tile

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kNumWStages; ++i) {
      mbar_init(bar(kBarWFull + i), 1);
      mbar_init(bar(kBarWEmpty + i), 1);
    }
    for (int i = 0; i < kAChunks; ++i) {
      mbar_init(bar(kBarAFull + i), HALF ? kEpiWarps / 2 : kEpiWarps);
      mbar_init(bar(kBarAEmpty + i), 1);
      mbar_init(bar(kBarASaved + i), 1);
      mbar_init(bar(kBarA0Sub + i), kEpiWarps);
    }
    mbar_init(bar(kBarAccFull), 1);
    if constexpr (PAIR) {
      for (int i = 0; i < kNumWStages; ++i) mbar_init(bar(kBarWPeer + i), 1);
      for (int i = 0; i < (HALF ? 16 : 8); ++i) mbar_init(bar(kBarAMma + i), HALF ? kEpiWarps : 2 * kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
      tmem_relinquish();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 4) {
  if (warp == 0) {
    // ======================= weight producer =======================
    if (lane == 0) {
      const uint8_t* head_images =
          p.packed + (int64_t)(p.input_kind == kInputX ? kImgHeadNatural : kImgHeadFused) * kWImageBytes;
      const uint8_t* body_images = p.packed + (int64_t)(kIsBwd ? kImgBodyT : kImgBody) * kWImageBytes;
      uint32_t it = 0;
      long long t_wait = 0;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        const int tile = tile_of(pt); (void)tile;
        for (int i = 0; i < kImagesPerTile; ++i, ++it) {
          const uint32_t ws = it % kNumWStages, ph = (it / kNumWStages) & 1u;
          const long long t0 = p.stats ? clock64() : 0;
          mbar_wait(bar(kBarWEmpty + ws), ph ^ 1u);
          if (p.stats) t_wait += clock64() - t0;
          mbar_arrive_expect_tx(bar(kBarWFull + ws), kWStageBytes);
          const uint8_t* src;
          if (kIsBwd) src = body_images + (int64_t)i * kWImageBytes;
          else src = i < 32 ? head_images + (int64_t)i * kWImageBytes : body_images + (int64_t)(i - 32) * kWImageBytes;
          if constexpr (HALF) {
            // my rows of the image = output features of K chunks rank and rank + 2 of the next layer: with the 2x2
            // accumulator layout the leader's rows land in TMEM lanes 0..63 and the peer's in lanes 64..127, so either
            // lane half ends up holding one EARLY (0 / 1) and one LATE (2 / 3) chunk of the next layer's operand, and the
            // chunks become ready in the order the GEMM takes them - the same order as in the other forms
            const uint32_t dst = smem_base + kSmemW + ws * kWStageBytes;
            bulk_g2s(dst, src + rank * 8192u, 8192u, bar(kBarWFull + ws));
            bulk_g2s(dst + 8192u, src + 16384u + rank * 8192u, 8192u, bar(kBarWFull + ws));
          } else {
            bulk_g2s(smem_base + kSmemW + ws * kWStageBytes, src + rank * kWStageBytes, kWStageBytes, bar(kBarWFull + ws));
          }
        }
      }
      if (p.stats) p.stats[blockIdx.x * 8 + 3] = t_wait;
    }
  } else if (PAIR && warp == 1 && rank == 1) {
    // ======================= pair form, peer CTA: forward "my weight half landed" to the leader =======================
    if (lane == 0) {
      uint32_t it = 0;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        for (int i = 0; i < kImagesPerTile; ++i, ++it) {
          const uint32_t ws = it % kNumWStages;
          mbar_wait(bar(kBarWFull + ws), (it / kNumWStages) & 1u);
          mbar_arrive_cluster_relaxed(mapa_cluster(bar(kBarWPeer + ws), 0));
        }
      }
    }
  } else if (HALF && warp == 1) {
    // ======================= MMA issuer, half form (leader CTA) =======================
    // At 64 cycles per MMA the ~100-cycle latency of a barrier probe would dominate a thread that waits for one barrier
    // after the other.  So the whole warp probes: lane i < 16 watches the operand barrier of (slot i / 4, k-step i % 4),
    // lanes 16..23 my weight stages, lanes 24..31 the peer's; one ballot tells the warp everything that has become ready.
    // Unlike the other forms this one never accumulates onto the residual stream (see kTmemY): results agree with theirs to
    // rounding, not bit for bit.
    static_assert(!HALF || kNumWStages == 8, "one group of four K chunks = one turn of the weight ring");
    constexpr uint32_t idesc = umma_idesc_f16(128, 256, 0, 0);
    uint32_t group = 0;        // groups of four K chunks (= 8 weight images = 16 operand units) issued so far
    uint32_t ready = 0;        // barriers of the current group known to have completed (bit = lane that watches it)
    const uint32_t my_bar = lane < 16 ? bar(kBarAMma + lane) : lane < 24 ? bar(kBarWFull + lane - 16) : bar(kBarWPeer + lane - 24);

    auto need = [&](uint32_t mask) {
      uint32_t spins = 0;
      while ((ready & mask) != mask) {
        const bool ok = ((ready >> lane) & 1u) || mbar_test_wait(my_bar, group & 1u);
        ready = __ballot_sync(0xffffffffu, ok);
        if (++spins > R2L_SPIN_LIMIT) __trap();
      }
      __syncwarp();   // orders the issuing lane behind the acquire of whichever lane saw the phase complete
    };
    for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
      for (int l = 0; l < kLayers; ++l) {
        const bool head = !kIsBwd && l == 0;
        const bool to_h = kIsBwd ? ((l & 1) == 0) : ((l & 1) != 0);
        // every GEMM starts from a zeroed accumulator: H (hidden layer / its gradient) or Y (joins the stream in the epilogue)
        const uint32_t d = tmem_base + (to_h ? kTmemH : kTmemY);
        const bool tr = p.trace != nullptr && pt == pair_id;
        for (int g = 0; g < (head ? kSamples / 4 : 1); ++g, ++group) {
          ready = 0;
          // Issue order inside a K chunk: first its SMALL split terms (A_hi W_lo, A_lo W_hi: ~2^-11 of the result), k-step by
          // k-step as the epilogue publishes them, then the four A_hi W_hi instructions.  After the last k-step of a layer is
          // published only 6 instructions remain to be issued (12 with the large term first).
#pragma unroll
          for (uint32_t c = 0; c < 4; ++c) {   // slot; stage 2c holds W_hi, stage 2c + 1 W_lo of this K chunk
            const uint32_t a_hi = smem_base + kSmemA + c * kSlotBytes, a_lo = a_hi + kPlane;
            const uint32_t b_hi = smem_base + kSmemW + (2 * c) * kWStageBytes, b_lo = b_hi + kWStageBytes;
            const uint32_t w_both = (3u << (16 + 2 * c)) | (3u << (24 + 2 * c));   // my and the peer's halves of both stages
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks) {
              need((1u << (4 * c + ks)) | w_both);
              // TMEM hazards (the epilogue's tcgen05.ld / st of earlier layers vs this layer's accumulator writes) are
              // all behind the first operand barrier of a layer: one tcgen05 fence per group is enough
              if (c == 0 && ks == 0) tc_fence_after_sync();
              if (elect_one_sync()) {
                if (tr && g == 0 && c == 0 && ks == 0) p.trace[((int64_t)blockIdx.x * 5 + 0) * 96 + l] = clock64();
                umma_f16_pair(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b_lo + 32 * ks, 16, 1024), idesc,
                              (g == 0 && c == 0 && ks == 0) ? 0u : 1u);
                umma_f16_pair(d, umma_desc_sw128(a_lo + 32 * ks, 16, 1024), umma_desc_sw128(b_hi + 32 * ks, 16, 1024), idesc, 1u);
              }
            }
            if (elect_one_sync()) {
              umma_commit_pair(bar(kBarWEmpty + 2 * c + 1));       // W_lo of this chunk has been read
#pragma unroll
              for (uint32_t ks = 0; ks < 4; ++ks)
                umma_f16_pair(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b_hi + 32 * ks, 16, 1024), idesc, 1u);
              umma_commit_pair(bar(kBarWEmpty + 2 * c));
              if (head) umma_commit_pair(bar(kBarAEmpty + c));   // the head's A chunks recycle through the 4 slots
            }
          }
          __syncwarp();
        }
        if (elect_one_sync()) {
          umma_commit_pair(bar(kBarAccFull));
          if (tr) p.trace[((int64_t)blockIdx.x * 5 + 1) * 96 + l] = clock64();
        }
        __syncwarp();
      }
    }
    if (p.stats && lane == 0) p.stats[blockIdx.x * 8 + 5] = global_timer_ns();
  } else if (warp == 1) {
    // ======================= MMA issuer (pair form: leader CTA only) =======================
    // The whole warp walks the schedule and waits; one elected lane issues.  (Issuing from inside `if (lane == 0)` makes
    // ptxas wrap every tcgen05.mma in an "any active thread" loop with per-instruction register -> uniform-register moves,
    // about 100 cycles per MMA; under elect.sync the descriptors stay in uniform registers.)
    {
      constexpr uint32_t idesc = umma_idesc_f16(PAIR ? 256 : 128, 256, 0, 0);   // pair: M = 256 over both CTAs
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
        if constexpr (PAIR) umma_f16_pair(d, a, b, idesc, acc); else umma_f16(d, a, b, idesc, acc);
      };
      auto commit = [&](uint32_t barrier) {   // pair: arrives on the barrier at this offset in BOTH CTAs
        if constexpr (PAIR) umma_commit_pair(barrier); else umma_commit(barrier);
      };
      uint32_t it = 0, a_phase = 0;
      long long t_a_head = 0, t_a_body = 0, t_w = 0;
      const long long t_begin = p.stats ? clock64() : 0;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        const int tile = tile_of(pt); (void)tile;
        for (int l = 0; l < kLayers; ++l) {
          // forward: l = 0 head (16 chunks, -> Z fresh), odd l -> H fresh, even l -> Z accumulate
          // backward: j = l: even j (da = g W2) -> H fresh, odd j (g += dh W1) -> Z accumulate
          const int nkc = (!kIsBwd && l == 0) ? kSamples : kAChunks;
          const bool to_h = kIsBwd ? ((l & 1) == 0) : ((l & 1) != 0);
          const bool fresh = to_h || (!kIsBwd && l == 0);
          const bool ring = !kIsBwd && l == 0;   // head: A chunks recycle through the 4 slots
          const uint32_t d = tmem_base + (to_h ? kTmemH : kTmemZ);
          const bool tr = p.trace != nullptr && pt == pair_id && lane == 0;
          for (int kk = 0; kk < nkc; ++kk) {
            const int kc = kk;
            const uint32_t slot = kc & 3;
            const uint32_t a_hi = smem_base + kSmemA + slot * kSlotBytes;
            const uint32_t a_lo = a_hi + kPlane;
            auto wait_w = [&](uint32_t i) {
              const long long t0 = p.stats ? clock64() : 0;
              mbar_wait(bar(kBarWFull + i % kNumWStages), (i / kNumWStages) & 1u);
              if constexpr (PAIR) mbar_wait(bar(kBarWPeer + i % kNumWStages), (i / kNumWStages) & 1u);   // the peer's half
              if (p.stats) t_w += clock64() - t0;
            };
            auto wait_a = [&](uint32_t barrier, uint32_t bit) {
              const long long t0 = p.stats ? clock64() : 0;
              // pair form: the 32-arrival MMA-facing twin of the barrier (bit 4+ks: k-step ks of slot 0; bit s: slot s)
              mbar_wait(bar(PAIR ? kBarAMma + (bit >= 4 ? bit - 4 : 4 + bit) : barrier), (a_phase >> bit) & 1u);
              if (p.stats) { if (l == 0) t_a_head += clock64() - t0; else t_a_body += clock64() - t0; }
              a_phase ^= 1u << bit;
              // one tcgen05 fence per layer (see the half-form issuer): everything the epilogue did to TMEM in earlier
              // layers is ordered before the first operand barrier of this layer
              if (kk == 0 && bit == 4) tc_fence_after_sync();
            };
            if (slot == 0) {
              // k-step granular: the hi-image MMAs of a k-step are issued as soon as its 16 columns are published;
              // the W_hi slot is released after them, the lo-image MMAs follow (same issue order as the other chunks)
              wait_w(it);
              const uint32_t b_hi = smem_base + kSmemW + (it % kNumWStages) * kWStageBytes;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                wait_a(kBarA0Sub + ks, 4 + ks);
                if (tr && kk == 0 && ks == 0) p.trace[((int64_t)blockIdx.x * 5 + 0) * 96 + l] = clock64();
                if (elect_one_sync()) {
                  mma(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b_hi + 32 * ks, 16, 1024), (fresh && kk == 0 && ks == 0) ? 0u : 1u);
                  mma(d, umma_desc_sw128(a_lo + 32 * ks, 16, 1024), umma_desc_sw128(b_hi + 32 * ks, 16, 1024), 1u);
                  if (ks == 3) commit(bar(kBarWEmpty + it % kNumWStages));
                }
              }
              ++it;
              wait_w(it);
              const uint32_t b_lo = smem_base + kSmemW + (it % kNumWStages) * kWStageBytes;
              if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  mma(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b_lo + 32 * ks, 16, 1024), 1u);
                commit(bar(kBarWEmpty + it % kNumWStages));
                if (ring) commit(bar(kBarAEmpty + slot));
              }
              ++it;
            } else {
              wait_a(kBarAFull + slot, slot);
              {  // W_hi image: A_hi*W_hi + A_lo*W_hi
                wait_w(it);
                const uint32_t b = smem_base + kSmemW + (it % kNumWStages) * kWStageBytes;
                if (elect_one_sync()) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    mma(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b + 32 * ks, 16, 1024), 1u);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    mma(d, umma_desc_sw128(a_lo + 32 * ks, 16, 1024), umma_desc_sw128(b + 32 * ks, 16, 1024), 1u);
                  commit(bar(kBarWEmpty + it % kNumWStages));
                }
                ++it;
              }
              {  // W_lo image: A_hi*W_lo
                wait_w(it);
                const uint32_t b = smem_base + kSmemW + (it % kNumWStages) * kWStageBytes;
                if (elect_one_sync()) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    mma(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b + 32 * ks, 16, 1024), 1u);
                  commit(bar(kBarWEmpty + it % kNumWStages));
                  if (ring) commit(bar(kBarAEmpty + slot));
                }
                ++it;
              }
            }
          }
          if (elect_one_sync()) commit(bar(kBarAccFull));
          if (tr) p.trace[((int64_t)blockIdx.x * 5 + 1) * 96 + l] = clock64();
        }
        if constexpr (kIsBwd) {
          // the last backward epilogue publishes 4 more chunks (d head pre-activation, consumed only by the
          // store warp): step over those phases so the parity bookkeeping stays aligned for the next tile
          for (uint32_t ks = 0; ks < 4; ++ks) {
            mbar_wait(bar(PAIR ? kBarAMma + ks : kBarA0Sub + ks), (a_phase >> (4 + ks)) & 1u);
            a_phase ^= 1u << (4 + ks);
          }
          for (uint32_t slot = 1; slot < kAChunks; ++slot) {
            mbar_wait(bar(PAIR ? kBarAMma + 4 + slot : kBarAFull + slot), (a_phase >> slot) & 1u);
            a_phase ^= 1u << slot;
          }
        }
      }
      if (p.stats && lane == 0) {
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
      // A chunk's slot is released (kBarASaved) once its store has finished READING shared memory: this warp waits for
      // exactly that (bulk wait_group.read 0, a few hundred cycles for 16-32 KiB) right after issuing the store - it has
      // nothing else to do until the next chunk is published - so every release follows its publish promptly and depends
      // on nothing younger.  The epilogue rewrites a slot a whole layer later and pays its (already satisfied) waits
      // while the layer's MMAs run.
      uint32_t a_phase = 0;
      int64_t issued = 0;
      const bool signal = kIsBwd && p.ready != nullptr;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        const int tile = tile_of(pt); (void)tile;
        uint8_t* dst = p.saved + (int64_t)tile * kSavedChunksPerTile * kAChunkBytes + row0 * 128u;
        int signalled = 0;   // operand groups (4 chunks = one layer's dY) of this tile already announced
        for (int i = 0; i < kSavedChunksPerTile; ++i, ++issued) {
          const uint32_t slot = i & 3;   // first chunks cycle the ring; body chunk c lives in slot c
          if (slot == 0 && !HALF) {
            // k-steps 0/2 and 1/3 of slot 0 are written by different warps: the chunk is complete when the last
            // k-step of both groups has been published
            mbar_wait(bar(kBarA0Sub + 2), a_phase & 1u);
            mbar_wait(bar(kBarA0Sub + 3), a_phase & 1u);
          } else {
            mbar_wait(bar(kBarAFull + slot), (a_phase >> slot) & 1u);
          }
          a_phase ^= 1u << slot;
          if constexpr (HALF) {   // my 64 rows of the hi and of the lo plane of the 128-row image
            bulk_s2g(dst + (int64_t)i * kAChunkBytes, smem_base + kSmemA + slot * kSlotBytes, kPlane);
            bulk_s2g(dst + (int64_t)i * kAChunkBytes + kPlaneBytes, smem_base + kSmemA + slot * kSlotBytes + kPlane, kPlane);
          } else {
            bulk_s2g(dst + (int64_t)i * kAChunkBytes, smem_base + kSmemA + slot * kSlotBytes, kAChunkBytes);
          }
          bulk_commit();
          bulk_wait_read<0>();           // this store has read its slot
          mbar_arrive(bar(kBarASaved + slot));
          if (signal && tile < p.num_tiles && slot == 3) {
            // everything but the 4 stores just issued has landed in global memory: announce those groups so the
            // weight-gradient kernel (running concurrently on idle SMs) may start on their layers
            bulk_wait_all<4>();
            for (; signalled < (i >> 2); ++signalled) flag_release_add(p.ready + signalled);
          }
        }
        if (signal && tile < p.num_tiles) {   // a dummy tile (pair form, odd tile count) must not be counted
          bulk_wait_all<0>();
          for (; signalled < kSavedChunksPerTile / 4; ++signalled) flag_release_add(p.ready + signalled);
        }
      }
      // drain: wait for the writes to land
      bulk_wait_all<0>();
    }
  }
  } else {
    // ======================= encoder / epilogue =======================
    const uint32_t ew = warp - 4;
    const uint32_t q = ew & 3u;       // TMEM lane quarter (== warp % 4)
    const uint32_t qt = ew >> 2;      // which quarter of a 64-column chunk this thread works on (0..3)
    // half form (2x2 TMEM layout, weight rows split as in the producer): TMEM lanes 64 hh .. 64 hh + 63, hh = q / 2, hold my
    // 64 rays' output features of K chunks hh (columns 0..63) and hh + 2 (columns 64..127) of the next layer
    const uint32_t hh = HALF ? (q >> 1) : 0u;
    const uint32_t row = HALF ? (q & 1u) * 32u + lane : q * 32u + lane;   // my ray inside this CTA's rows
    constexpr uint32_t kMyChunks = HALF ? 2 : 4;                     // K chunks this thread produces: c = hh + 2 cc / c = cc
    auto my_chunk = [&](uint32_t cc) { return HALF ? hh + 2u * cc : cc; };
    auto mine = [&](uint32_t c) { return !HALF || (c & 1u) == hh; };
    auto tmem_col = [&](uint32_t c) { return HALF ? 64u * (c >> 1) : 64u * c; };   // where chunk c's 64 columns start
    const uint32_t tmem_row = tmem_base + ((q * 32u) << 16);
    const float* cumbias = reinterpret_cast<const float*>(p.packed + kPackOffCumBias);
    const float* headb = reinterpret_cast<const float*>(p.packed + kPackOffHeadB);
    const float* b1 = reinterpret_cast<const float*>(p.packed + kPackOffB1);
    const float* b2 = reinterpret_cast<const float*>(p.packed + kPackOffB2);
    const float* tailw = reinterpret_cast<const float*>(p.packed + kPackOffTailW);
    const float* tailb = reinterpret_cast<const float*>(p.packed + kPackOffTailB);
    float* hrow = p.scratch + ((int64_t)blockIdx.x * kTileM + row) * kWidth;
    uint32_t acc_phase = 0;
    uint32_t saved_phase = 0;   // per-slot parity of kBarASaved
    (void)cumbias; (void)headb; (void)b1; (void)b2; (void)tailb; (void)tail_smem;
    // Epilogue column ownership inside a 64-column chunk: k-steps g0 = qt>>1 and g0+2, and inside each k-step the
    // 8-column unit u = qt&1, i.e. the 16-byte operand units 2g+u.  K-steps 0/2 belong to the warps with qt in {0,1},
    // k-steps 1/3 to qt in {2,3}: the first 16 columns of a layer's output are ready after 8 warps did 8 columns each.
    const uint32_t g0 = qt >> 1, uu = qt & 1u;

    // kBarA0Sub[ks] counts all 16 warps: owners arrive when their part of k-step ks is written, the others at once.
    // single form: the CTA-local barriers serve the MMA thread and the store warp.  pair form: the MMA-facing barriers
    // live in the leader (local arrive there, signal-only remote arrive from the peer); the local ones are only needed
    // by each CTA's own store warp, i.e. in the training modes.
    auto arrive_mma = [&](uint32_t idx) {
      if (lane == 0) {
        if (rank == 0) mbar_arrive(bar(kBarAMma + idx));
        else mbar_arrive_cluster_relaxed(mapa_cluster(bar(kBarAMma + idx), 0));
      }
    };
    auto arrive_sub = [&](uint32_t ks) {
      if constexpr (PAIR) arrive_mma(ks);
      if ((!PAIR || kSave) && lane == 0) mbar_arrive(bar(kBarA0Sub + ks));
    };
    auto arrive_chunk = [&](uint32_t slot) {    // chunk `slot` (1..3) complete as far as this warp is concerned
      if constexpr (PAIR) arrive_mma(4 + slot);
      if ((!PAIR || kSave) && lane == 0) mbar_arrive(bar(kBarAFull + slot));
    };
    // half form: every slot is handed to the MMA issuer k-step by k-step (kBarAMma[4 slot + ks], the 8 warps per CTA that
    // own the slot's feature half arrive: writers of the k-step when it is written, the others at once), and as a whole to
    // this CTA's store warp (kBarAFull[slot], training modes)
    auto arrive_unit = [&](uint32_t slot, uint32_t ks) { arrive_mma(4 * slot + ks); };
    auto arrive_saved_chunk = [&](uint32_t slot) {
      if (kSave && lane == 0) mbar_arrive(bar(kBarAFull + slot));
    };
    auto make_visible = [&]() {   // generic-proxy smem writes -> tensor core (async proxy), TMEM reads ordered
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
    };
    auto publish = [&](uint32_t slot) {          // this warp's part of a whole 64-column chunk is written
      make_visible();
      if constexpr (HALF) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) arrive_unit(slot, ks);
        arrive_saved_chunk(slot);
      } else if (slot == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) arrive_sub(ks);
      } else {
        arrive_chunk(slot);
      }
    };
    // before rewriting a slot in save modes: the store warp must have copied the previous content out
    // (half form: a slot always has the same writers, the warps of feature half slot / 2)
    auto wait_saved = [&](uint32_t slot, bool first_use) {
      if (kSave) {
        if (!first_use) mbar_wait(bar(kBarASaved + slot), (saved_phase >> slot) & 1u);
        if (!first_use) saved_phase ^= 1u << slot;
      }
    };

    bool first_tile = true;
    for (int pt = pair_id; pt < num_ptiles; pt += num_pairs, first_tile = false) {
      const int tile = tile_of(pt);
      const int64_t grow = (int64_t)tile * kTileM + row0 + row;
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
        } else if (p.input_kind == kInputRays9 && valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            o[c] = __ldg(p.in0 + grow * 9 + c);
            dd[c] = __ldg(p.in0 + grow * 9 + 3 + c);
          }
        } else if (p.input_kind == kInputPose && valid) {
          // PointSampler.__init__ / sample_test (nerf_raybased.py:80-86, :94-99): dirs = [(i - W/2)/f, -(j - H/2)/f, -1],
          // rays_d[r] = sum_k dirs[k] c2w[r][k] (summed left to right), rays_o = c2w[:, 3]
          const int64_t frame = (int64_t)p.img_h * p.img_w;
          const int64_t pose = grow / frame;
          const int pix = (int)(grow - pose * frame);
          const int pj = pix / p.img_w, pi = pix - pj * p.img_w;
          const float* m = p.in0 + pose * 12;
          const float dx = __fdiv_rn(__fsub_rn((float)pi, (float)p.img_w * .5f), p.focal);
          const float dy = -__fdiv_rn(__fsub_rn((float)pj, (float)p.img_h * .5f), p.focal);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            dd[c] = __fadd_rn(__fadd_rn(__fmul_rn(dx, __ldg(m + 4 * c)), __fmul_rn(dy, __ldg(m + 4 * c + 1))), -__ldg(m + 4 * c + 2));
            o[c] = __ldg(m + 4 * c + 3);
          }
        }
        for (int c = 0; c < kSamples; ++c) {
          const uint32_t slot = c & 3;
          // half form: the eight threads of a ray split the samples by slot parity (as the chunks of the body layers)
          const bool writer = mine(slot);
          float f[16];
          if (!writer) {
            continue;
          } else if (p.input_kind == kInputX) {
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
          const uint32_t chunk_addr = smem_base + kSmemA + slot * kSlotBytes;
          store_a_unit(chunk_addr, row, 2 * qt, &f[0], kPlane);
          store_a_unit(chunk_addr, row, 2 * qt + 1, &f[8], kPlane);
          publish(slot);
        }
      } else {
        // ---- backward prologue: d logit = d rgb * rgb (1 - rgb);  g = d logit . W_tail  (dL/dz_43 = dL/dh_skip) ----
        // The whole backward runs on loss_scale * dL/d(.) (a power of two chosen per call from max |d rgb| by
        // r2l_bwd_prep_kernel, so that the fp16 dY planes sit in the normal range); dw.cu divides it out again.
        const float loss_scale = p.bwd_scale ? __ldg(p.bwd_scale) : 1.f;
        float dl[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float y = __ldg(p.rgb_in + grow * 3 + c);
            dl[c] = __ldg(p.grad_rgb + grow * 3 + c) * y * (1.f - y) * loss_scale;
          }
        }
        for (int c = 0; c < kAChunks; ++c) {
          if (!mine(c)) continue;   // half form: the other feature half's warps write this chunk
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
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(HALF ? v[i] : v[i] * kWeightScale);   // pair / single: Z in accumulator units
          tmem_st8(tmem_row + kTmemZ + tmem_col(c) + 16u * qt, &r[0]);
          tmem_st8(tmem_row + kTmemZ + tmem_col(c) + 16u * qt + 8, &r[8]);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(hrow + col)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          tmem_st_wait();
          wait_saved(c, first_tile);
          const uint32_t chunk_addr = smem_base + kSmemA + c * kSlotBytes;
          store_a_unit(chunk_addr, row, 2 * qt, &v[0], kPlane);
          store_a_unit(chunk_addr, row, 2 * qt + 1, &v[8], kPlane);
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
          bias = l == 0 ? headb : ((l & 1) ? b1 + (l >> 1) * kWidth : (HALF ? b2 + ((l >> 1) - 1) * kWidth : cumbias + (l >> 1) * kWidth));
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
        // side data of a chunk (bias / ReLU mask of my two 8-column units); the first chunk's is fetched while the MMAs run.
        // (Fetching further ahead does not pay: the proxy fence of every publish waits for all loads in flight.)
        float4 bq[4];
        uint4 mq[2];
        auto load_side = [&](uint32_t c) {
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
                mq[h] = __ldg(reinterpret_cast<const uint4*>(plane + (row0 + row) * 128u + (((2u * (g0 + 2 * h) + uu) ^ (row & 7u)) << 4)));
            }
          }
        };
        load_side(my_chunk(0));
        const bool feeds_mma = !last;                     // the last epilogue of a tile produces no further GEMM input
        const bool produces_chunk = feeds_mma || kIsBwd;  // backward's last output (d head pre-activation) is saved for dw
        // The slots I am about to rewrite were released right after the previous epilogue published them (store warp):
        // those waits are paid here, while this layer's MMAs run, not between accumulator-complete and the publishes.
        if (produces_chunk) {
#pragma unroll
          for (uint32_t cc = 0; cc < kMyChunks; ++cc) wait_saved(my_chunk(cc), false);
        }
        // half form: the GEMM result sits in H or Y; when it joins the stream (!from_h) the stream's current value is read
        // from Z beside it, the sum is formed in fp32 and written back to Z (the head writes h there).  No MMA touches Z, so
        // the first chunk's stream values are fetched while the layer's MMAs still run.
        const bool joins = HALF && !from_h;
        const bool has_old = joins && (kIsBwd || l > 0);
        // accumulator -> value: 1 / kWeightScale, times (1 + eps) in the half form to undo in expectation what the tensor
        // core's truncating accumulation took from the sum (ChainParams::inv_body / inv_head)
        const float inv = HALF ? ((!kIsBwd && l == 0) ? p.inv_head : p.inv_body) : kInvWeightScale;
        // The value this thread forms from an accumulator element r is ALWAYS fmaf(r, inv, bq): bq = bias (forward), 0
        // (backward), plus the stream value when the result joins the stream - added into bq here, before the accumulator
        // wait, so that the path from "accumulator complete" to the first publish is the same short code in every layer
        // ((z + b2) + y instead of the reference's (y + b2) + z: one fp32 rounding each, either way).
        auto add_stream = [&](uint32_t cc) {      // bq += Z[my two units of chunk cc]; warp-uniform branch at the call
          uint32_t zo[16];
          const uint32_t tz = tmem_row + kTmemZ + 64u * cc + 8u * uu;
          tmem_ld8(tz + 16u * g0, &zo[0]);
          tmem_ld8(tz + 16u * (g0 + 2), &zo[8]);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            bq[2 * h].x += __uint_as_float(zo[8 * h + 0]); bq[2 * h].y += __uint_as_float(zo[8 * h + 1]);
            bq[2 * h].z += __uint_as_float(zo[8 * h + 2]); bq[2 * h].w += __uint_as_float(zo[8 * h + 3]);
            bq[2 * h + 1].x += __uint_as_float(zo[8 * h + 4]); bq[2 * h + 1].y += __uint_as_float(zo[8 * h + 5]);
            bq[2 * h + 1].z += __uint_as_float(zo[8 * h + 6]); bq[2 * h + 1].w += __uint_as_float(zo[8 * h + 7]);
          }
        };
        if constexpr (kIsBwd) {
#pragma unroll
          for (int i = 0; i < 4; ++i) bq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (has_old) add_stream(0);
        mbar_wait(bar(kBarAccFull), acc_phase);
        acc_phase ^= 1u;
        tc_fence_after_sync();
        const bool tr = p.trace != nullptr && pt == pair_id && warp == 4 && lane == 0;
        if (tr) p.trace[((int64_t)blockIdx.x * 5 + 2) * 96 + l] = clock64();
        for (uint32_t cc = 0; cc < kMyChunks; ++cc) {
          const uint32_t c = my_chunk(cc);
          uint32_t r[16];
          const uint32_t tacc = tmem_row + (from_h ? kTmemH : (HALF ? kTmemY : kTmemZ)) + 64u * cc + 8u * uu;
          tmem_ld8(tacc + 16u * g0, &r[0]);
          tmem_ld8(tacc + 16u * (g0 + 2), &r[8]);
          if (cc > 0) {
            load_side(c);
            if constexpr (kIsBwd) {
#pragma unroll
              for (int i = 0; i < 4; ++i) bq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (has_old) add_stream(cc);   // (its wait also covers the two accumulator loads just issued)
          }
          if (produces_chunk) {
            if constexpr (HALF) {
              if (feeds_mma) { arrive_unit(c, 1 - g0); arrive_unit(c, 3 - g0); }   // the k-steps of this chunk I do not write
            } else {
              if (c == 0) { arrive_sub(1 - g0); arrive_sub(3 - g0); }   // the k-steps of chunk 0 I do not write
            }
          }
          tmem_ld_wait();
          const uint32_t chunk_addr = smem_base + kSmemA + c * kSlotBytes;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t g = g0 + 2 * h;
            const uint32_t col = 64u * c + 16u * g + 8u * uu;
            float v[8];
            if constexpr (!kIsBwd) {
              // the accumulators hold kWeightScale * (a . W): the rescale is an exact power of two inside the bias add
              v[0] = fmaf(__uint_as_float(r[8 * h + 0]), inv, bq[2 * h].x);
              v[1] = fmaf(__uint_as_float(r[8 * h + 1]), inv, bq[2 * h].y);
              v[2] = fmaf(__uint_as_float(r[8 * h + 2]), inv, bq[2 * h].z);
              v[3] = fmaf(__uint_as_float(r[8 * h + 3]), inv, bq[2 * h].w);
              v[4] = fmaf(__uint_as_float(r[8 * h + 4]), inv, bq[2 * h + 1].x);
              v[5] = fmaf(__uint_as_float(r[8 * h + 5]), inv, bq[2 * h + 1].y);
              v[6] = fmaf(__uint_as_float(r[8 * h + 6]), inv, bq[2 * h + 1].z);
              v[7] = fmaf(__uint_as_float(r[8 * h + 7]), inv, bq[2 * h + 1].w);
              if (relu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
              }
              if (l == 0) {   // z_0 = h is kept for the outer skip (:543)
                reinterpret_cast<float4*>(hrow + col)[0] = make_float4(v[0], v[1], v[2], v[3]);
                reinterpret_cast<float4*>(hrow + col)[1] = make_float4(v[4], v[5], v[6], v[7]);
              }
              if constexpr (!HALF) if (l == 0) {   // pair / single: seed the in-place residual stream, in accumulator units
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(v[i] * kWeightScale);
                tmem_st8(tmem_row + kTmemZ + tmem_col(c) + 16u * g + 8u * uu, w);
                tmem_st_wait();
              }
            } else {
#pragma unroll
              // g += dh W1 when the result joins the stream (bq holds g), else bq = 0
              v[0] = fmaf(__uint_as_float(r[8 * h + 0]), inv, bq[2 * h].x);
              v[1] = fmaf(__uint_as_float(r[8 * h + 1]), inv, bq[2 * h].y);
              v[2] = fmaf(__uint_as_float(r[8 * h + 2]), inv, bq[2 * h].z);
              v[3] = fmaf(__uint_as_float(r[8 * h + 3]), inv, bq[2 * h].w);
              v[4] = fmaf(__uint_as_float(r[8 * h + 4]), inv, bq[2 * h + 1].x);
              v[5] = fmaf(__uint_as_float(r[8 * h + 5]), inv, bq[2 * h + 1].y);
              v[6] = fmaf(__uint_as_float(r[8 * h + 6]), inv, bq[2 * h + 1].z);
              v[7] = fmaf(__uint_as_float(r[8 * h + 7]), inv, bq[2 * h + 1].w);
              if (last) {  // + dL/dz_43 through the outer skip
                const float4 s0 = reinterpret_cast<const float4*>(hrow + col)[0];
                const float4 s1 = reinterpret_cast<const float4*>(hrow + col)[1];
                v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
                v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
              }
              if (masked) {
                // ReLU mask from the hi plane of the saved forward operand (a > 0  <=>  fp16 hi != 0 for a >= 2^-25; a >= 0 always)
                const uint32_t w[4] = {mq[h].x, mq[h].y, mq[h].z, mq[h].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if ((w[e] & 0x00007FFFu) == 0u) v[2 * e] = 0.f;
                  if ((w[e] & 0x7FFF0000u) == 0u) v[2 * e + 1] = 0.f;
                }
              }
            }
            if (produces_chunk) {
              store_a_unit(chunk_addr, row, 2 * g + uu, v, kPlane);
              if constexpr (HALF) {
                make_visible();
                if (feeds_mma) arrive_unit(c, g);
                if (tr && c == 0 && g == 0) p.trace[((int64_t)blockIdx.x * 5 + 3) * 96 + l] = clock64();
              } else if (c == 0) {
                make_visible();
                arrive_sub(g);
                if (tr && g == 0) p.trace[((int64_t)blockIdx.x * 5 + 3) * 96 + l] = clock64();
              }
            }
            if constexpr (HALF) {
              if (joins) {   // the stream's new value goes back to Z behind the publish (warp-uniform branch)
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(v[i]);
                tmem_st8(tmem_row + kTmemZ + 64u * cc + 16u * g + 8u * uu, w);
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
          if constexpr (HALF) {
            if (produces_chunk) arrive_saved_chunk(c);   // both of my k-steps were made visible above
          } else {
            if (produces_chunk && c > 0) publish(c);
          }
          if constexpr (!kIsBwd) {
            if (last) {
              // park this chunk's partial tail dots (the same 16 columns in every form) where one thread per ray can add
              // all sixteen in a fixed order: the accumulator region H, idle in the last layer, is shared by the four
              // threads of a ray; the half form's eight threads per ray sit in two lane halves and use shared memory
              if constexpr (HALF) {
                float* tp = tail_smem + ((row * 4u + c) * 4u + qt) * 3u;
                tp[0] = dot0; tp[1] = dot1; tp[2] = dot2;
              } else {
                uint32_t w4[4] = {__float_as_uint(dot0), __float_as_uint(dot1), __float_as_uint(dot2), 0u};
                tmem_st4(tmem_row + kTmemH + 16u * c + 4u * qt, w4);
              }
              dot0 = dot1 = dot2 = 0.f;
            }
          }
        }
        if (joins) tmem_st_wait();   // the stream values written above are read back by this thread two layers on
        if (tr) p.trace[((int64_t)blockIdx.x * 5 + 4) * 96 + l] = clock64();
      }
      if constexpr (!kIsBwd) {
        // rgb = sigmoid(sum of the 16 partial dots in the order (chunk 0: quarter 0..3), (chunk 1: ...), ... + bias)
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        if constexpr (HALF) {
          named_bar_sync(1, kEpiWarps * 32);
          if (hh == 0 && qt == 0) {
            const float* tp = tail_smem + row * 48u;
#pragma unroll
            for (int j = 0; j < 16; ++j) { s0 += tp[3 * j]; s1 += tp[3 * j + 1]; s2 += tp[3 * j + 2]; }
          }
        } else {
          tmem_st_wait();
          tc_fence_before_sync();
          named_bar_sync(1, kEpiWarps * 32);
          tc_fence_after_sync();
          if (qt == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t a[16];
              tmem_ld16(tmem_row + kTmemH + 16u * c, a);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                s0 += __uint_as_float(a[4 * j]); s1 += __uint_as_float(a[4 * j + 1]); s2 += __uint_as_float(a[4 * j + 2]);
              }
            }
          }
        }
        if (hh == 0 && qt == 0 && valid) {
          const float y0 = 1.f / (1.f + expf(-(s0 + __ldg(tailb + 0))));
          const float y1 = 1.f / (1.f + expf(-(s1 + __ldg(tailb + 1))));
          const float y2 = 1.f / (1.f + expf(-(s2 + __ldg(tailb + 2))));
          if (p.rgb != nullptr) {
            p.rgb[grow * 3 + 0] = y0;
            p.rgb[grow * 3 + 1] = y1;
            p.rgb[grow * 3 + 2] = y2;
          }
          if (p.rgb8 != nullptr) {   // to8b (nerf_raybased.py:16): (255 * clip(x, 0, 1)).astype(uint8), i.e. truncation
            p.rgb8[grow * 3 + 0] = (uint8_t)__fmul_rn(255.f, fminf(fmaxf(y0, 0.f), 1.f));
            p.rgb8[grow * 3 + 1] = (uint8_t)__fmul_rn(255.f, fminf(fmaxf(y1, 0.f), 1.f));
            p.rgb8[grow * 3 + 2] = (uint8_t)__fmul_rn(255.f, fminf(fmaxf(y2, 0.f), 1.f));
          }
        }
        if constexpr (HALF) named_bar_sync(1, kEpiWarps * 32);   // the partials are read before the next tile overwrites them
        tc_fence_before_sync();
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the leader's MMAs touch the peer's smem / TMEM until the very end
  tc_fence_after_sync();
  if (warp == 2) {
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------------------------
// Single-layer self test: C[128,256] = A[128,256] * W^T using exactly the operand images, descriptors
// and TMEM read-back the chain kernel uses. `images` = 8 consecutive 32 KiB images (4 chunks x {hi,lo}).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) r2l_umma_selftest_kernel(const float* __restrict__ A,
                                                                  const uint8_t* __restrict__ images,
                                                                  float* __restrict__ C) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t wbase = smem_base + kABytes;
  const uint32_t bar_w = smem_base + kABytes + 2 * kWImageBytes;
  const uint32_t bar_m = bar_w + 8;
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kABytes + 2 * kWImageBytes + 16);
  const int warp = threadIdx.x >> 5;
  const uint32_t row = threadIdx.x;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 256);
    tmem_relinquish();
  }
  // A operand: every thread splits its own row
  for (int c = 0; c < kAChunks; ++c) {
    for (uint32_t unit = 0; unit < 8; ++unit) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = A[row * kWidth + 64 * c + 8 * unit + i];
      store_a_unit(smem_base + c * kAChunkBytes, row, unit, v);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, 256, 0, 0);
    for (int kc = 0; kc < kAChunks; ++kc) {
      mbar_arrive_expect_tx(bar_w, 2 * kWImageBytes);
      bulk_g2s(wbase, images + (int64_t)(2 * kc) * kWImageBytes, kWImageBytes, bar_w);
      bulk_g2s(wbase + kWImageBytes, images + (int64_t)(2 * kc + 1) * kWImageBytes, kWImageBytes, bar_w);
      mbar_wait(bar_w, kc & 1);
      tc_fence_after_sync();
      const uint32_t a_hi = smem_base + kc * kAChunkBytes, a_lo = a_hi + kPlaneBytes;
      for (int ks = 0; ks < 4; ++ks)
        umma_f16(tmem_base, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(wbase + 32 * ks, 16, 1024),
                  idesc, (kc == 0 && ks == 0) ? 0u : 1u);
      for (int ks = 0; ks < 4; ++ks)
        umma_f16(tmem_base, umma_desc_sw128(a_lo + 32 * ks, 16, 1024), umma_desc_sw128(wbase + 32 * ks, 16, 1024),
                  idesc, 1u);
      for (int ks = 0; ks < 4; ++ks)
        umma_f16(tmem_base, umma_desc_sw128(a_hi + 32 * ks, 16, 1024),
                  umma_desc_sw128(wbase + kWImageBytes + 32 * ks, 16, 1024), idesc, 1u);
      umma_commit(bar_m);
      mbar_wait(bar_m, kc & 1);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  for (int c8 = 0; c8 < 8; ++c8) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32 * c8, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) C[row * kWidth + 32 * c8 + i] = __uint_as_float(r[i]) * kInvWeightScale;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ----------------------------------------------------------------------------------------------
// Micro-benchmark (debug): the tensor pipe alone.  Each CTA issues `reps` body-layer MMA patterns (4 chunks x
// (A_hi W_hi, A_lo W_hi, A_hi W_lo) x 4 k-steps = 48 instructions of M128 N256 K16) on resident operands, no TMA,
// no epilogue, and reports the cycles.  6144 cycles per layer = the 8192 FLOP/cycle/SM peak.
// ----------------------------------------------------------------------------------------------
template <int FORM>
__global__ void __launch_bounds__(128, 1) r2l_mma_rate_kernel(int reps, int variant, long long* __restrict__ out) {
  constexpr bool PAIR = FORM != kFormSingle;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t wbase = smem_base + kABytes;
  const uint32_t bar_m = smem_base + kABytes + 2 * kWImageBytes;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kABytes + 2 * kWImageBytes + 16);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // variant bit 0: non-zero operands; bits 1..: issue pattern (half form: 1 = the chain kernel's addresses and order,
  // 2 = its addresses, hi / lo products grouped per chunk)
  const uint32_t fill = (variant & 1) ? 0x3F803F80u : 0u;
  const int pattern = variant >> 1;
  for (uint32_t i = threadIdx.x; i < (kABytes + 2 * kWImageBytes) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(fill, fill, fill, fill);
  if (threadIdx.x == 0) {
    mbar_init(bar_m, 1);
    mbar_init(bar_m + 8, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    if constexpr (PAIR) { tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 256); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 256); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0 && rank == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(FORM == kFormPair ? 256 : 128, 256, 0, 0);
    auto mma = [&](uint64_t a, uint64_t b) {
      if constexpr (PAIR) umma_f16_pair(tmem_base, a, b, idesc, 1u); else umma_f16(tmem_base, a, b, idesc, 1u);
    };
    const long long t0 = clock64();
    if (FORM == kFormHalf && pattern != 0) {
      constexpr ChainGeom G = chain_geom(kFormHalf);
      auto mma_d = [&](uint32_t d, uint64_t a, uint64_t b) { umma_f16_pair(d, a, b, idesc, 1u); };
      for (int r = 0; r < reps; ++r) {
        const uint32_t d = tmem_base + ((r & 1) ? G.tmem_h : 0u);
        for (int pr = 0; pr < 2; ++pr) {
          for (int hk = 0; hk < 2; ++hk)
            for (int s2 = 0; s2 < 2; ++s2) {
              const uint32_t c = pr + 2 * s2, st_hi = 4 * pr + s2;
              const uint32_t a_hi = smem_base + c * G.slot, a_lo = a_hi + G.plane, b_hi = smem_base + G.off_w + st_hi * G.w_stage_bytes;
              if (pattern == 1 || pattern == 3) {
                for (int k2 = 0; k2 < 2; ++k2) {
                  const uint32_t ks = 2 * hk + k2;
                  if (pattern == 3) tc_fence_after_sync();   // what a tcgen05 fence in front of every MMA pair costs
                  mma_d(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b_hi + 32 * ks, 16, 1024));
                  mma_d(d, umma_desc_sw128(a_lo + 32 * ks, 16, 1024), umma_desc_sw128(b_hi + 32 * ks, 16, 1024));
                }
              } else {
                for (int k2 = 0; k2 < 2; ++k2) mma_d(d, umma_desc_sw128(a_hi + 32 * (2 * hk + k2), 16, 1024), umma_desc_sw128(b_hi + 32 * (2 * hk + k2), 16, 1024));
                for (int k2 = 0; k2 < 2; ++k2) mma_d(d, umma_desc_sw128(a_lo + 32 * (2 * hk + k2), 16, 1024), umma_desc_sw128(b_hi + 32 * (2 * hk + k2), 16, 1024));
              }
              if (hk == 1) umma_commit_pair(bar_m + 8);
            }
          for (int s2 = 0; s2 < 2; ++s2) {
            const uint32_t c = pr + 2 * s2, st_lo = 4 * pr + 2 + s2;
            const uint32_t a_hi = smem_base + c * G.slot, b_lo = smem_base + G.off_w + st_lo * G.w_stage_bytes;
            for (int ks = 0; ks < 4; ++ks) mma_d(d, umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(b_lo + 32 * ks, 16, 1024));
            umma_commit_pair(bar_m + 8);
          }
        }
      }
    } else
    for (int r = 0; r < reps; ++r) {
      for (int kc = 0; kc < kAChunks; ++kc) {
        const uint32_t a_hi = smem_base + kc * kAChunkBytes, a_lo = a_hi + kPlaneBytes;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma(umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(wbase + 32 * ks, 16, 1024));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma(umma_desc_sw128(a_lo + 32 * ks, 16, 1024), umma_desc_sw128(wbase + 32 * ks, 16, 1024));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma(umma_desc_sw128(a_hi + 32 * ks, 16, 1024), umma_desc_sw128(wbase + kWImageBytes + 32 * ks, 16, 1024));
      }
    }
    if constexpr (PAIR) umma_commit_pair(bar_m); else umma_commit(bar_m);
    mbar_wait(bar_m, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 0) {
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 256); else tmem_dealloc(tmem_base, 256);
  }
}

template <int FORM>
static cudaError_t launch_mma_rate_form(int reps, int variant, int grid, long long* out, cudaStream_t stream) {
  const int smem = kABytes + 2 * kWImageBytes + 64 + 1024;
  cudaError_t e = cudaFuncSetAttribute(r2l_mma_rate_kernel<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = FORM == kFormSingle ? 1 : 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, r2l_mma_rate_kernel<FORM>, reps, variant, out);
}

// form: the chain kernel's launch forms (single: M128 cta_group::1; pair: M256 cta_group::2; half: M128 cta_group::2);
// pair / half: grid even, out[2i] holds the cycles of pair i
cudaError_t launch_mma_rate(int form, int variant, int reps, int grid, long long* out, cudaStream_t stream) {
  switch (form) {
    case kFormSingle: return launch_mma_rate_form<kFormSingle>(reps, variant, grid, out, stream);
    case kFormPair: return launch_mma_rate_form<kFormPair>(reps, variant, grid, out, stream);
    case kFormHalf: return launch_mma_rate_form<kFormHalf>(reps, variant, grid, out, stream);
  }
  return cudaErrorInvalidValue;
}

template <int MODE, int FORM>
static cudaError_t launch_chain_mode(const ChainParams& p, int grid, cudaStream_t stream) {
  constexpr uint32_t smem = chain_smem_bytes(FORM);
  cudaError_t e = cudaFuncSetAttribute(r2l_chain_kernel<MODE, FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);          // pair / half form: even, CTAs 2i and 2i+1 form a cluster
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = FORM == kFormSingle ? 1 : 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, r2l_chain_kernel<MODE, FORM>, p);
}

template <int MODE>
static cudaError_t launch_chain_form(int form, const ChainParams& p, int grid, cudaStream_t stream) {
  switch (form) {
    case kFormSingle: return launch_chain_mode<MODE, kFormSingle>(p, grid, stream);
    case kFormPair: return launch_chain_mode<MODE, kFormPair>(p, grid, stream);
    case kFormHalf: return launch_chain_mode<MODE, kFormHalf>(p, grid, stream);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_chain(int mode, int form, const ChainParams& p, int grid, cudaStream_t stream) {
  switch (mode) {
    case kFwdInfer: return launch_chain_form<kFwdInfer>(form, p, grid, stream);
    case kFwdTrain: return launch_chain_form<kFwdTrain>(form, p, grid, stream);
    case kBwd: return launch_chain_form<kBwd>(form, p, grid, stream);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_umma_selftest(const float* A, const void* images, float* C, cudaStream_t stream) {
  const int smem = kABytes + 2 * kWImageBytes + 64 + 1024;
  cudaError_t e = cudaFuncSetAttribute(r2l_umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  r2l_umma_selftest_kernel<<<1, 128, smem, stream>>>(A, static_cast<const uint8_t*>(images), C);
  return cudaGetLastError();
}

}  // namespace r2l
