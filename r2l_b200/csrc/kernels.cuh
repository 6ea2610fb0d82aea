// Host-visible launch interface of the CUDA kernels (internal to the library; the public ABI is include/r2l_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.cuh"

namespace r2l {

enum ChainMode : int { kFwdInfer = 0, kFwdTrain = 1, kBwd = 2 };

struct ChainParams {
  // forward inputs
  const float* in0;       // rays_o[N,3] | pts[N,48] | x[N,1008] | c2w[P,3,4] | rays9[N,9]
  const float* in1;       // rays_d[N,3] | unused
  const float* t_rand;    // [N,16] or nullptr (kInputRays only)
  float z_lo[kSamples];   // z_vals (no jitter) or `lower` (jitter)
  float z_diff[kSamples]; // `upper - lower` (jitter)
  const uint8_t* packed;
  float* rgb;             // forward out [N,3] (may be nullptr when rgb8 is given)
  uint8_t* rgb8;          // optional forward out [N,3] uint8 = to8b(rgb) (nerf_raybased.py:16); nullptr = off
  int img_h, img_w;       // kInputPose: frame size; pixel (i = column, j = row) of ray r is r % (H W)
  float focal;            // kInputPose: focal length in pixels
  float inv_body, inv_head;  // half form: accumulator -> value factors 1 / kWeightScale * (1 + eps), see c_api.cu: g_debias
  float* scratch;         // [gridDim.x][128][256] fp32: head output (fwd) / dL/dz_43 (bwd) for the outer skip
  long long* stats;       // optional [gridDim.x][8] cycle counters (debug), nullptr in production
  long long* trace;       // optional [gridDim.x][5][96] clock64 stamps of the first tile's layers (debug)
  // training
  uint8_t* saved;         // out: operand images this pass stores, [tile][chunk][32 KiB]
  float* zf_out;          // kFwdTrain out: z_43 + h, [N,256] fp32
  const uint8_t* fwd_saved;  // kBwd in: the forward pass's `saved`
  const float* rgb_in;    // kBwd in: forward rgb [N,3]
  const float* grad_rgb;  // kBwd in: dL/d rgb [N,3]
  const float* bwd_scale; // kBwd in (device): [0] = loss scale the backward runs on (power of two), [1] = its inverse
  int* ready;             // kBwd out (optional): [87] counters, ready[g] += 1 when this CTA's tile has stored operand group g
                          // (g = 0: dL/dz_43, g = 1 + j: output of backward epilogue j); lets dw.cu run concurrently
  int64_t n_rays;
  int num_tiles;
  int input_kind;
};

// saved operand images per 128-ray tile (32 KiB chunks)
constexpr int kFwdSavedChunks = kSamples + 4 * kBodyLayers;   // 16 PE chunks + input of each body Linear
constexpr int kBwdSavedChunks = kAChunks + 4 * kBodyLayers;   // dL/dz_43 + output gradient of each body Linear (+ head)

struct DwParams {
  const uint8_t* fwd_saved;
  const uint8_t* bwd_saved;
  float* grads;           // flat [kNumParams]; weight blocks are overwritten, (or accumulated when `accumulate`)
  int num_tiles;
  int input_kind;         // kInputX: head features in natural order, else fused-PE order
  int accumulate;
  // schedule: unit u (in release order: body layer 85 .. 0, then the 4 head column groups) is cut into unit_splits[u]
  // ray-tile ranges, one work item each, items unit_first[u] .. unit_first[u] + unit_splits[u] - 1.  A piece of a split unit
  // writes its partial result to partials[item]; the last piece of the unit to finish (ticket) sums them in order.
  uint8_t unit_splits[kBodyLayers + 4];
  uint16_t unit_first[kBodyLayers + 4];
  int deterministic;      // split units: 1 = partials in scratch + ordered sum by the last piece, 0 = L2 reductions into grads
  int num_items;          // work items = pieces of all units, numbered in release order (item = "CTA" above)
  int item_lo, item_hi;   // the items THIS launch processes (one launch per gradient chunk, see r2l_backward_chunked)
  int grid;               // persistent CTAs (<= SMs); they claim items in order from `queue`
  int* queue;             // zeroed counter
  float* partials;        // [num_items][256*256 + 256] scratch (only slots of split units are used)
  int* tickets;           // [90] zeroed counters (only needed when some unit is split)
  long long* times;       // optional debug: [unit][4] globaltimer stamps (start, flag seen, MMAs done, end)
  const float* bwd_scale; // device: [1] = 1 / (loss scale of the dY operands); results are multiplied by it (exact)
  const int* ready;       // optional: wait until ready[group of this unit] == ready_target before streaming (see ChainParams)
  int ready_target;       // store warps that announce each group: num_tiles (x 2 when the chain ran in its half form)
};

struct TailGradParams {
  const float* zf;        // [N,256]
  const float* rgb;       // [N,3]
  const float* grad_rgb;  // [N,3]
  float* grads;           // flat; tail.0.weight / tail.0.bias are overwritten
  float* partials;        // [256][771] scratch
  int* ticket;            // zeroed counter (left zero)
  int64_t n_rays;
};
constexpr size_t kTailPartialBytes = (size_t)256 * 771 * sizeof(float);

struct TeacherParams {
  const float* pts;          // [P,3] sample points (ray-major: point p belongs to ray p / samples_per_ray)
  const float* viewdirs;     // [P / samples_per_ray, 3] normalised view directions
  const float* rays_o;       // alternative to pts: rays_o[R,3], rays_d[R,3], z_vals[R,samples_per_ray]; the point is built in-kernel
  const float* rays_d;
  const float* z_vals;
  const float* x_embedded;   // alternative input: [P,90] already embedded (pts 63 | views 27); pts/viewdirs unused
  const uint8_t* packed;
  float* raw;                // [P,4] = (rgb, sigma) before any activation
  int64_t n_points;
  int64_t samples_per_ray;
  int num_tiles;
};

cudaError_t launch_teacher_pack(const float* params, void* packed, cudaStream_t stream);
cudaError_t launch_teacher(const TeacherParams& p, int grid, cudaStream_t stream);
cudaError_t launch_pack(const float* params, void* packed, cudaStream_t stream);
enum : int { kFormSingle = 0, kFormPair = 1, kFormHalf = 2 };   // launch forms of the chain kernels, see chain.cu
cudaError_t launch_chain(int mode, int form, const ChainParams& p, int grid, cudaStream_t stream);   // pair / half: grid even, clusters of 2
cudaError_t launch_dw(const DwParams& p, cudaStream_t stream);
cudaError_t launch_tail_grads(const TailGradParams& p, cudaStream_t stream);
// zeroes the `ready_ints` flag words at `ready` and writes {S, 1/S} to scale_out: S = 2^k with S * max |grad_rgb| in [2^9, 2^10)
cudaError_t launch_bwd_prep(const float* grad_rgb, int64_t n_values, int* ready, int ready_ints, float* scale_out, cudaStream_t stream);
cudaError_t launch_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int64_t n_rays, int n_samples,
                               int white_bkgd, float* rgb_map, float* disp_map, float* acc_map, float* weights,
                               float* depth_map, cudaStream_t stream);
cudaError_t launch_sample_pdf_merge(const float* z_vals, const float* weights, const float* u, int64_t u_stride, int64_t n_rays,
                                    int S, int M, float* z_samples, float* z_merged, const float* bins_in, cudaStream_t stream);
cudaError_t launch_embed(const float* x, float* out, int64_t n, int dim, int L, int style, cudaStream_t stream);
constexpr int kDpMaxWorld = 16;
struct DpParams {           // dp.cu: reduce-scatter + Adam + all-gather over peer memory
  float* grads[kDpMaxWorld];     // every rank's gradient buffer (mine and the peers' IPC mappings)
  float* params[kDpMaxWorld];    // every rank's parameter buffer
  uint32_t* flags[kDpMaxWorld];  // every rank's flag block: [0,16) "gradient of rank s complete", [16,32) "slice of rank s written"
  uint32_t* state;               // local, per launch slot: [0] epochs completed, [1] CTA completion counter
  int flag_base;                 // first flag word of this launch slot (32 words per slot)
  float* exp_avg;                // local full-size moment buffers; only [shard_lo, shard_hi) is used on this rank
  float* exp_avg_sq;
  const float* hyper;            // device: {lr / bias_correction1, 1 / sqrt(bias_correction2)} (r2l_adam_schedule_dev)
  int64_t shard_lo, shard_hi;
  float w1, beta2, w2, eps;
  int rank, world;
  int variant;                   // DEBUG timing switches (r2l_debug_set_dp_grid's second argument), 0 in production
};
cudaError_t launch_dp_adam(const DpParams& p, int grid, cudaStream_t stream);
// pool.cu: hard-example ray pool (main.py:1325-1347, :1410-1425)
cudaError_t launch_pool_draw(const float* pool_rows, const int* state, int n_out, uint64_t seed, const long long* counters,
                             float* dst_rows, int* slots_out, cudaStream_t stream);
uint32_t pool_slot_host(uint32_t j, uint32_t size, uint64_t seed, uint64_t step);
cudaError_t launch_pool_update(const float* rays9, const float* err, int n, int k, float* pool_rows, int* state, const int* slots_out,
                               int* picked, cudaStream_t stream);
struct AdamSchedule {
  double lrate, warmup_start_lr, warmup_end, decay_rate, decay_steps, beta1, beta2;   // warmup_end = 0: no warm-up
};
cudaError_t launch_adam_schedule(const AdamSchedule& sc, long long* counters, float* hyper, cudaStream_t stream);
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float w1, float beta2, float w2, float eps,
                        float step_size, float inv_bc2_sqrt, const float* hyper, cudaStream_t stream);
cudaError_t launch_mse_loss_grad(const float* rgb, const float* target, int64_t n, int target_stride, float grad_scale, float loss_scale,
                                 float* grad_rgb, float* per_ray, float* loss, float* scratch, cudaStream_t stream);
}  // namespace r2l
#include <string>
namespace r2l {
std::string read_ray_shards(const char* const* paths, int n_paths, float* dst, int64_t floats_per_shard, int n_threads);   // io.cu (host)
cudaError_t launch_mma_rate(int form, int variant, int reps, int grid, long long* out, cudaStream_t stream);
cudaError_t launch_umma_selftest(const float* A, const void* images, float* C, cudaStream_t stream);

}  // namespace r2l
