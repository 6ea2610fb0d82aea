// Geometry of the R2L W256/D88 light-field network and of every buffer the kernels exchange.
// Reference architecture: /root/reference/model/nerf_raybased.py:483-544 (NeRF_v3_2), :443-465 (ResMLP).
#pragma once
#include <stdint.h>

namespace r2l {

// ---- network geometry (the README configuration, reference README.md:51) ----
constexpr int kWidth = 256;          // --netwidth
constexpr int kBlocks = 43;          // (88 - 2) / 2 ResMLP blocks, nerf_raybased.py:515-518
constexpr int kBodyLayers = 86;      // 2 Linear per block
constexpr int kSamples = 16;         // --n_sample_per_ray
constexpr int kFreqs = 10;           // --multires
constexpr int kEmbed = 2 * kFreqs + 1;              // 21, PositionalEmbedder.embed_dim (:196)
constexpr int kInDim = kSamples * 3 * kEmbed;       // 1008
constexpr int kInDimPad = 1024;                     // 16 K-chunks of 64
constexpr int kOutDim = 3;

// ---- flat fp32 parameter buffer, in state_dict order ----
// head.0.weight [256,1008], head.0.bias [256],
// body.k.body.0.weight [256,256], body.k.body.0.bias [256], body.k.body.2.weight, body.k.body.2.bias (k=0..42),
// tail.0.weight [3,256], tail.0.bias [3]
constexpr int64_t kOffHeadW = 0;
constexpr int64_t kOffHeadB = (int64_t)kWidth * kInDim;               // 258048
constexpr int64_t kOffBody = kOffHeadB + kWidth;                      // 258304
constexpr int64_t kLinearStride = (int64_t)kWidth * kWidth + kWidth;  // 65792
constexpr int64_t kBlockStride = 2 * kLinearStride;                   // 131584
constexpr int64_t kOffTailW = kOffBody + kBlocks * kBlockStride;      // 5916416
constexpr int64_t kOffTailB = kOffTailW + kOutDim * kWidth;           // 5917184
constexpr int64_t kNumParams = kOffTailB + kOutDim;                   // 5917187
// body layer l in [0,86): linear index l (block l/2, inner l%2)
__host__ __device__ constexpr int64_t off_body_w(int l) { return kOffBody + (int64_t)l * kLinearStride; }
__host__ __device__ constexpr int64_t off_body_b(int l) { return off_body_w(l) + (int64_t)kWidth * kWidth; }

// ---- operand format ----
// Every tensor-core operand plane in this library is IEEE fp16.  A value is carried as two planes x ~= hi + lo (ptx.cuh:
// split2): 22 significant bits while |x| >= 2^-3, an absolute floor of 2^-25 below (fp16 subnormals).  Weights are therefore
// packed pre-multiplied by kWeightScale (their lo planes would otherwise sit in the subnormal range: |w| ~ 2^-5 at default
// init) and every accumulator read multiplies by 1 / kWeightScale, an exact power of two.
constexpr float kWeightScale = 64.f;
constexpr float kInvWeightScale = 1.f / kWeightScale;

// ---- tiles ----
constexpr int kTileM = 128;    // rays per CTA tile = UMMA M = TMEM lanes
constexpr int kChunkK = 64;    // K elements per 128-byte swizzled row
constexpr int kPlaneBytes = kTileM * 128;        // one fp16 plane (hi or lo) of a [128 x 64] A chunk
constexpr int kAChunkBytes = 2 * kPlaneBytes;    // hi + lo
constexpr int kAChunks = 4;                      // 256 / 64
constexpr int kABytes = kAChunks * kAChunkBytes; // 131072: whole [128 x 256] A operand, hi+lo
constexpr int kWImageBytes = kWidth * 128;       // 32768: [256 n x 64 k] fp16, one plane

// ---- packed weight stream (fp16 planes, UMMA K-major SWIZZLE_128B images of 32 KiB) ----
// image order = consumption order of the forward chain kernel:
//   head (fused-PE feature order) : 16 chunks x {hi, lo}            images [0, 32)
//   head (natural feature order)  : 16 chunks x {hi, lo}            images [32, 64)
//   body layer l = 0..85          : 4 chunks x {hi, lo}             images [64 + 8 l, 64 + 8 l + 8)
// then the transposed body weights in the order the backward chain consumes them:
//   for k = 42..0: W2_k^T (4 x {hi,lo}), W1_k^T (4 x {hi,lo})       images [752, 752 + 688)
constexpr int kImgHeadFused = 0;
constexpr int kImgHeadNatural = 32;
constexpr int kImgBody = 64;
constexpr int kImgBodyT = kImgBody + 8 * kBodyLayers;   // 752
constexpr int kNumImages = kImgBodyT + 8 * kBodyLayers; // 1440
constexpr int64_t kPackedImagesBytes = (int64_t)kNumImages * kWImageBytes;  // 47,185,920
// fp32 side tables appended after the images
//   cumbias[44][256] : cumbias[k] = sum_{j<k} b2_j  (the residual stream lives un-biased in TMEM)
//   headb[256], b1[43][256], tailw[3][256], tailb[4], b2[43][256]
constexpr int64_t kPackOffCumBias = kPackedImagesBytes;
constexpr int64_t kPackOffHeadB = kPackOffCumBias + 44 * kWidth * 4;
constexpr int64_t kPackOffB1 = kPackOffHeadB + kWidth * 4;
constexpr int64_t kPackOffTailW = kPackOffB1 + kBlocks * kWidth * 4;
constexpr int64_t kPackOffTailB = kPackOffTailW + kOutDim * kWidth * 4;
constexpr int64_t kPackOffB2 = kPackOffTailB + 16;                       // b2[43][256]: second-layer biases as they are
constexpr int64_t kPackedBytes = kPackOffB2 + kBlocks * kWidth * 4;

// Fused-PE feature order inside K-chunk s (= sample point s of the ray), 64 slots:
//   pair p = c*10 + f (coordinate c, frequency f): slot 2p = sin(x_c 2^f), slot 2p+1 = cos(x_c 2^f)
//   slots 60,61,62 = x_0,x_1,x_2 ; slot 63 = 0
// and the reference feature index of each (PositionalEmbedder.__call__, nerf_raybased.py:198-208:
// per coordinate j: [sin f0..f9, cos f0..f9, x]).
__host__ __device__ inline int fused_slot_to_feature(int s, int slot) {
  if (slot < 60) {
    const int p = slot >> 1, c = p / kFreqs, f = p % kFreqs;
    return (3 * s + c) * kEmbed + ((slot & 1) ? kFreqs + f : f);
  }
  if (slot < 63) return (3 * s + (slot - 60)) * kEmbed + 2 * kFreqs;
  return -1;
}

// byte offset of element (row, k) inside one [rows x 64] 16-bit K-major SWIZZLE_128B plane
__host__ __device__ inline uint32_t sw128_offset(uint32_t row, uint32_t k) {
  return row * 128u + ((((k >> 3) ^ (row & 7u)) << 4) | ((k & 7u) << 1));
}

// ---- teacher NeRF (D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], use_viewdirs) ----
constexpr int kTeacherLayers = 10;            // tensor-core layers T0..T9 (see teacher.cu)
constexpr int kTeacherImagePairs = 39;        // K chunks over all layers: 1+4+4+4+4+5+4+4+4+5
constexpr int64_t kTeacherNumParams = 595844; // floats in the teacher state_dict
constexpr int64_t kTeacherPackOffTables = (int64_t)2 * kTeacherImagePairs * kWImageBytes;
constexpr int64_t kTeacherPackedBytes = kTeacherPackOffTables + (kTeacherLayers * kWidth + kWidth + 3 * 128 + 4) * 4;

enum InputKind : int {
  kInputRays = 0,  // rays_o[N,3], rays_d[N,3] (+ optional t_rand[N,16]); points and PE built in-kernel
  kInputPts = 1,   // pts[N,48] as returned by PointSampler.sample_*; PE built in-kernel
  kInputX = 2,     // x[N,1008] materialised PositionalEmbedder output
  kInputRays9 = 4, // rays9[N,9] = (o | d | rgb) rows of a ray shard (utils/create_data.py:820-872): o, d read in place at stride 9
  kInputPose = 3,  // c2w[P,3,4] camera poses; ray r = pixel (r % (H W)) of pose r / (H W): PointSampler.sample_test in-kernel
};

}  // namespace r2l
