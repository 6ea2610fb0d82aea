// extern "C" boundary (include/r2l_b200.h). Argument checking, launch-geometry policy, error strings.
#include <cstdio>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/r2l_b200.h"
#include "kernels.cuh"

namespace {
thread_local char g_err[512] = "";
long long* g_stats = nullptr;  // debug cycle counters, see r2l_debug_set_stats
long long* g_trace = nullptr;  // debug time stamps, see r2l_debug_set_trace
int g_form = -1;               // launch form of the chain kernels (chain.cu): 0 single, 1 pair, 2 half, -1 = chosen per call

// Debias of the tensor core's truncating accumulation (half form; chain.cu, DESIGN.md "Precision"): every tcgen05.mma
// rounds its fp32 accumulator toward zero, so a fresh 48-instruction GEMM comes out short by a few ulp on average.  The
// epilogue multiplies the accumulator by (1 + eps): eps_body for the K = 256 GEMMs, eps_head for the K = 1024 head.
// The values are calibrated on the device (tools/gpu_accum_calibrate.py; profiles/r2_summary.md) and can be overridden
// for such measurements with r2l_debug_set_accum_debias.
// Calibration (profiles/r2_summary.md, gpurun_out/r2_03/calibrate.log): the mean signed forward error crosses zero at
// eps_body = 12..14 x 2^-24 on lego-pose and on stress rays alike; rms relative RGB error 1.4e-6 -> 0.65e-6 (lego),
// 1.8e-6 -> 1.0e-6 (stress).  The head's optimum is flat between 32 and 64 x 2^-24.
float g_debias[2] = {12.f / 16777216.f, 32.f / 16777216.f};
void set_debias(r2l::ChainParams& p) {
  p.inv_body = r2l::kInvWeightScale * (1.f + g_debias[0]);
  p.inv_head = r2l::kInvWeightScale * (1.f + g_debias[1]);
}

int fail(const char* fmt, const char* detail) {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return -1;
}
int check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -2;
}
// kernels this library has launched (or recorded into a CUDA graph being captured) since the last reset: bench.py counts
// the launches of one iteration with it instead of asserting a constant
long long g_launches = 0;
int check_launch(cudaError_t e, const char* what) {
  if (e == cudaSuccess) ++g_launches;
  return check(e, what);
}
int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}
// side stream + events per device for the concurrent dW launch of r2l_backward
constexpr int kMaxGradChunks = 8;
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaEvent_t chunk_done[kMaxGradChunks] = {};   // r2l_backward_chunked: gradient chunk i of the last call is complete
};
std::mutex g_side_mutex;
SideStream g_side[64];
SideStream* side_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_side_mutex);
  SideStream& s = g_side[dev];
  if (!s.stream) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (int i = 0; i < kMaxGradChunks; ++i)
      if (cudaEventCreateWithFlags(&s.chunk_done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &s;
}
constexpr size_t kReadyBytes = 1024;  // 87 readiness counters + 90 reduction tickets (ints) + queue + loss scale {S, 1/S}, padded
constexpr int kDwUnits = r2l::kBodyLayers + 4;
constexpr int kDwMaxCtas = 720;       // scratch slots for partial weight gradients (one per CTA of a split unit)
constexpr size_t kDwPartialBytes = (size_t)kDwMaxCtas * (256 * 256 + 256) * sizeof(float);
// Schedule of the weight-gradient kernel (dw.cu; see DwParams).  Units are numbered in the order the backward chain
// releases them.  g_dw_sched: {t1, t2, t3, s_after} for the concurrent mode (units < t1 whole, < t2 in 2 pieces, < t3 in
// 4, < t4 in 8, the rest in 16) and the piece count of every unit in the serial mode; negative = built-in default.
int g_dw_sched[5] = {-1, -1, -1, -1, -1};
int g_deterministic = 0;          // r2l_set_deterministic
void dw_schedule(r2l::DwParams& d, bool concurrent) {
  // Concurrent mode: the kernel's persistent CTAs claim pieces in release order and follow the chain one piece behind,
  // so the piece length sets how long dW keeps running after the chain has released its last layers: 4 pieces per unit
  // (8 ray tiles of a 4096-ray batch, ~30 us), 8 for the last two body layers and 16 for the head's column groups, which
  // are released in the chain's last microseconds.
  const int t1 = g_dw_sched[0] >= 0 ? g_dw_sched[0] : 0, t2 = g_dw_sched[1] >= 0 ? g_dw_sched[1] : 0,
            t3 = g_dw_sched[2] >= 0 ? g_dw_sched[2] : 84, t4 = g_dw_sched[4] >= 0 ? g_dw_sched[4] : 86;
  // Serial mode (large batches, dW after the chain): whole units; the kernel is HBM-bound there (it streams 171 KiB per
  // ray) and more pieces only add scratch traffic (measured: 1, 2, 3, 4, 8 pieces at 18,944 and 98,304 rays)
  const int serial = g_dw_sched[3] > 0 ? g_dw_sched[3] : 1;
  int first = 0;
  for (int u = 0; u < kDwUnits; ++u) {
    int s = concurrent ? (u < t1 ? 1 : u < t2 ? 2 : u < t3 ? 4 : u < t4 ? 8 : 16) : serial;
    if (s > 16) s = 16;
    if (first + s + (kDwUnits - 1 - u) > kDwMaxCtas) s = 1;   // never more items than scratch slots (deterministic mode)
    while (s > 1 && 2 * s > d.num_tiles) --s;      // at least two ray tiles per piece
    d.unit_splits[u] = (uint8_t)s;
    d.unit_first[u] = (uint16_t)first;
    first += s;
  }
  d.num_items = first;
  const int sms = sm_count();
  d.grid = first < sms ? first : sms;
}

int num_tiles(int64_t n_rays) { return (int)((n_rays + r2l::kTileM - 1) / r2l::kTileM); }
// Which launch form a chain kernel takes.  Two SMs that share a tile (half form) finish it in 0.69 of the time (measured,
// 4096 rays: 0.321 vs 0.467 ms), so it wins whenever there are SM pairs to spare (tiles <= SMs / 2).  With more tiles
// than SM pairs every SM has work either way and a pair finishing two tiles (pair form) or an SM finishing one (single
// form) is the better use of it: measured at 160,000 rays 5.31 (half) / 4.25 (pair) / 4.52 ms (single), training kernels
// at 98,304 rays +15 % in the half form.  Large batches therefore run in the pair form.
int chain_form(int mode, int64_t n_rays) {
  if (g_form >= 0) return g_form;
  (void)mode;
  return num_tiles(n_rays) <= sm_count() / 2 ? r2l::kFormHalf : r2l::kFormPair;
}
int fwd_grid(int64_t n_rays, int mode) {
  const int sms = sm_count();
  const int t = num_tiles(n_rays);
  const int form = chain_form(mode, n_rays);
  if (form != r2l::kFormSingle) {   // CTA pairs: an even grid; pair form: one pair per two tiles, half form: one pair per tile
    const int pairs = form == r2l::kFormHalf ? t : (t + 1) / 2, max_pairs = sms / 2;
    return 2 * (pairs < max_pairs ? pairs : max_pairs);
  }
  return t < sms ? t : sms;
}
int plain_grid(int64_t n_units) {   // one CTA per 128-unit tile, at most one per SM (single-CTA kernels)
  const int sms = sm_count();
  const int t = num_tiles(n_units);
  return t < sms ? t : sms;
}
int even_tiles(int64_t n_rays) { return (num_tiles(n_rays) + 1) & ~1; }   // pair mode may run one dummy tile
cudaError_t launch_chain_any(int mode, const r2l::ChainParams& p, int grid, cudaStream_t stream) {
  return r2l::launch_chain(mode, chain_form(mode, p.n_rays), p, grid, stream);
}
}  // namespace

extern "C" {

const char* r2l_last_error(void) { return g_err; }
int r2l_abi_version(void) { return 1; }

size_t r2l_packed_bytes(void) { return (size_t)r2l::kPackedBytes; }

size_t r2l_fwd_workspace_bytes(int64_t n_rays) {
  // head-output scratch: one [128,256] fp32 tile per resident CTA (sized for the largest grid we launch)
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int t = 2 * num_tiles(n_rays);   // the half form runs two CTAs per tile, each with its own scratch rows
  const int g = t < sms ? t : sms;
  return (size_t)(g > 0 ? g : 1) * r2l::kTileM * r2l::kWidth * sizeof(float) + kReadyBytes;
}

int r2l_pack_weights(const float* params, void* packed, void* stream) {
  if (!params || !packed) return fail("r2l_pack_weights: %s", "null pointer");
  return check_launch(r2l::launch_pack(params, packed, (cudaStream_t)stream), "r2l_pack_weights");
}

namespace {
int fill_inputs(r2l::ChainParams& p, const char* who, int input_kind, const float* in0, const float* in1,
                const float* t_rand, const float* z_lo, const float* z_diff) {
  if (!in0) return fail("%s: null input pointer", who);
  const bool rays = input_kind == R2L_INPUT_RAYS || input_kind == R2L_INPUT_RAYS9;
  if ((input_kind < 0 || input_kind > 2) && input_kind != R2L_INPUT_RAYS9) return fail("%s: unknown input_kind", who);
  if (input_kind == R2L_INPUT_RAYS && !in1) return fail("%s: rays input needs in1", who);
  if (rays && !z_lo) return fail("%s: rays input needs z_lo", who);
  if (t_rand && (!rays || !z_diff)) return fail("%s: t_rand needs a rays input kind and z_diff", who);
  p.in0 = in0;
  p.in1 = in1;
  p.t_rand = t_rand;
  for (int i = 0; i < r2l::kSamples; ++i) {
    p.z_lo[i] = z_lo ? z_lo[i] : 0.f;
    p.z_diff[i] = z_diff ? z_diff[i] : 0.f;
  }
  p.input_kind = input_kind;
  return 0;
}
bool misaligned(const void* q) { return ((uintptr_t)q & 15) != 0; }
}  // namespace

int r2l_forward(int input_kind, const float* in0, const float* in1, const float* t_rand, const float* z_lo,
                const float* z_diff, const void* packed, float* rgb, void* workspace, size_t workspace_bytes,
                int64_t n_rays, void* stream) {
  if (n_rays == 0) return 0;
  if (n_rays < 0) return fail("r2l_forward: %s", "negative n_rays");
  if (!packed || !rgb || !workspace) return fail("r2l_forward: %s", "null pointer");
  if (workspace_bytes < r2l_fwd_workspace_bytes(n_rays)) return fail("r2l_forward: %s", "workspace too small");
  if (misaligned(packed) || misaligned(workspace)) return fail("r2l_forward: %s", "packed/workspace must be 16-byte aligned");
  r2l::ChainParams p;
  memset(&p, 0, sizeof(p));
  if (int rc = fill_inputs(p, "r2l_forward", input_kind, in0, in1, t_rand, z_lo, z_diff)) return rc;
  p.packed = static_cast<const uint8_t*>(packed);
  p.rgb = rgb;
  p.scratch = static_cast<float*>(workspace);
  p.n_rays = n_rays;
  p.num_tiles = num_tiles(n_rays);
  set_debias(p);
  p.stats = g_stats;
  p.trace = g_trace;
  return check_launch(launch_chain_any(r2l::kFwdInfer, p, fwd_grid(n_rays, r2l::kFwdInfer), (cudaStream_t)stream), "r2l_forward");
}

int r2l_render_poses(const float* c2w, int64_t n_poses, int height, int width, float focal, const float* z_vals,
                     const void* packed, float* rgb, uint8_t* rgb8, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_poses == 0) return 0;
  if (n_poses < 0 || height <= 0 || width <= 0) return fail("r2l_render_poses: %s", "bad frame geometry");
  if (!(focal > 0.f)) return fail("r2l_render_poses: %s", "focal must be positive");
  const int64_t n_rays = n_poses * (int64_t)height * width;
  if (!c2w || !z_vals || !packed || !workspace || (!rgb && !rgb8)) return fail("r2l_render_poses: %s", "null pointer");
  if (workspace_bytes < r2l_fwd_workspace_bytes(n_rays)) return fail("r2l_render_poses: %s", "workspace too small");
  if (misaligned(packed) || misaligned(workspace)) return fail("r2l_render_poses: %s", "packed/workspace must be 16-byte aligned");
  r2l::ChainParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = c2w;
  for (int i = 0; i < r2l::kSamples; ++i) p.z_lo[i] = z_vals[i];
  p.input_kind = r2l::kInputPose;
  p.img_h = height;
  p.img_w = width;
  p.focal = focal;
  p.packed = static_cast<const uint8_t*>(packed);
  p.rgb = rgb;
  p.rgb8 = rgb8;
  p.scratch = static_cast<float*>(workspace);
  p.n_rays = n_rays;
  p.num_tiles = num_tiles(n_rays);
  set_debias(p);
  p.stats = g_stats;
  p.trace = g_trace;
  return check_launch(launch_chain_any(r2l::kFwdInfer, p, fwd_grid(n_rays, r2l::kFwdInfer), (cudaStream_t)stream), "r2l_render_poses");
}

size_t r2l_bwd_workspace_bytes(int64_t n_rays) { return r2l_fwd_workspace_bytes(n_rays) + kDwPartialBytes + r2l::kTailPartialBytes; }

size_t r2l_train_fwd_saved_bytes(int64_t n_rays) {
  return (size_t)even_tiles(n_rays) * r2l::kFwdSavedChunks * r2l::kAChunkBytes;
}
size_t r2l_train_bwd_saved_bytes(int64_t n_rays) {
  return (size_t)even_tiles(n_rays) * r2l::kBwdSavedChunks * r2l::kAChunkBytes;
}

int r2l_forward_train(int input_kind, const float* in0, const float* in1, const float* t_rand, const float* z_lo,
                      const float* z_diff, const void* packed, float* rgb, float* zf, void* fwd_saved,
                      void* workspace, size_t workspace_bytes, int64_t n_rays, void* stream) {
  if (n_rays == 0) return 0;
  if (n_rays < 0) return fail("r2l_forward_train: %s", "negative n_rays");
  if (!packed || !rgb || !zf || !fwd_saved || !workspace) return fail("r2l_forward_train: %s", "null pointer");
  if (workspace_bytes < r2l_fwd_workspace_bytes(n_rays)) return fail("r2l_forward_train: %s", "workspace too small");
  if (misaligned(packed) || misaligned(workspace) || misaligned(fwd_saved) || misaligned(zf))
    return fail("r2l_forward_train: %s", "buffers must be 16-byte aligned");
  r2l::ChainParams p;
  memset(&p, 0, sizeof(p));
  if (int rc = fill_inputs(p, "r2l_forward_train", input_kind, in0, in1, t_rand, z_lo, z_diff)) return rc;
  p.packed = static_cast<const uint8_t*>(packed);
  p.rgb = rgb;
  p.zf_out = zf;
  p.saved = static_cast<uint8_t*>(fwd_saved);
  p.scratch = static_cast<float*>(workspace);
  p.n_rays = n_rays;
  p.num_tiles = num_tiles(n_rays);
  set_debias(p);
  p.stats = g_stats;
  p.trace = g_trace;
  return check_launch(launch_chain_any(r2l::kFwdTrain, p, fwd_grid(n_rays, r2l::kFwdTrain), (cudaStream_t)stream), "r2l_forward_train");
}

namespace {
// unit (release order of dw.cu: body layer 85 .. 0, then the head's four column groups) where gradient chunk `chunk`
// starts, for split layers given in descending order: chunk 0 = body layers >= split[0] (+ the tail), ...,
// last chunk = body layers < split[last] + the head
int chunk_first_unit(int n_chunks, const int* split_layers, int chunk) {
  if (chunk <= 0) return 0;
  if (chunk >= n_chunks) return kDwUnits;
  return r2l::kBodyLayers - split_layers[chunk - 1];
}
int check_splits(int n_chunks, const int* split_layers, const char* who) {
  if (n_chunks < 1 || n_chunks > kMaxGradChunks) return fail("%s: n_chunks out of range", who);
  if (n_chunks > 1 && !split_layers) return fail("%s: split_layers is null", who);
  for (int i = 0; i + 1 < n_chunks; ++i)
    if (split_layers[i] <= 0 || split_layers[i] >= r2l::kBodyLayers || (i > 0 && split_layers[i] >= split_layers[i - 1]))
      return fail("%s: split_layers must be strictly descending body-layer indices in (0, 86)", who);
  return 0;
}
int backward_impl(int input_kind, const void* packed, const float* rgb, const float* grad_rgb, const float* zf,
                  const void* fwd_saved, void* bwd_saved, float* grads, void* workspace, size_t workspace_bytes,
                  int64_t n_rays, void* stream, int n_chunks, const int* split_layers, int reserve_sms);
}  // namespace

int r2l_backward(int input_kind, const void* packed, const float* rgb, const float* grad_rgb, const float* zf,
                 const void* fwd_saved, void* bwd_saved, float* grads, void* workspace, size_t workspace_bytes,
                 int64_t n_rays, void* stream) {
  return backward_impl(input_kind, packed, rgb, grad_rgb, zf, fwd_saved, bwd_saved, grads, workspace, workspace_bytes, n_rays, stream,
                       1, nullptr, 0);
}

int r2l_backward_chunked(int input_kind, const void* packed, const float* rgb, const float* grad_rgb, const float* zf,
                         const void* fwd_saved, void* bwd_saved, float* grads, void* workspace, size_t workspace_bytes,
                         int64_t n_rays, void* stream, int n_chunks, const int* split_layers, int reserve_sms) {
  if (int rc = check_splits(n_chunks, split_layers, "r2l_backward_chunked")) return rc;
  if (reserve_sms < 0 || reserve_sms > 64) return fail("r2l_backward_chunked: %s", "reserve_sms out of range");
  return backward_impl(input_kind, packed, rgb, grad_rgb, zf, fwd_saved, bwd_saved, grads, workspace, workspace_bytes, n_rays, stream,
                       n_chunks, split_layers, reserve_sms);
}

int r2l_grad_chunk_range(int n_chunks, const int* split_layers, int chunk, int64_t* lo, int64_t* hi) {
  if (int rc = check_splits(n_chunks, split_layers, "r2l_grad_chunk_range")) return rc;
  if (chunk < 0 || chunk >= n_chunks || !lo || !hi) return fail("r2l_grad_chunk_range: %s", "bad chunk / null pointer");
  // flat buffer order: head, body 0 .. 85, tail; chunk 0 is the TOP of the buffer
  *hi = chunk == 0 ? r2l::kNumParams : r2l::off_body_w(split_layers[chunk - 1]);
  *lo = chunk == n_chunks - 1 ? 0 : r2l::off_body_w(split_layers[chunk]);
  return 0;
}

int r2l_stream_wait_grad_chunk(int chunk, void* stream) {
  if (chunk < 0 || chunk >= kMaxGradChunks) return fail("r2l_stream_wait_grad_chunk: %s", "bad chunk");
  SideStream* side = side_stream();
  if (!side) return fail("r2l_stream_wait_grad_chunk: %s", "no device state");
  return check(cudaStreamWaitEvent((cudaStream_t)stream, side->chunk_done[chunk], 0), "r2l_stream_wait_grad_chunk");
}

namespace {
int backward_impl(int input_kind, const void* packed, const float* rgb, const float* grad_rgb, const float* zf,
                  const void* fwd_saved, void* bwd_saved, float* grads, void* workspace, size_t workspace_bytes,
                  int64_t n_rays, void* stream, int n_chunks, const int* split_layers, int reserve_sms) {
  if (n_rays == 0) return 0;
  if (n_rays < 0) return fail("r2l_backward: %s", "negative n_rays");
  if (!packed || !rgb || !grad_rgb || !zf || !fwd_saved || !bwd_saved || !grads || !workspace)
    return fail("r2l_backward: %s", "null pointer");
  if ((input_kind < 0 || input_kind > 2) && input_kind != R2L_INPUT_RAYS9) return fail("r2l_backward: %s", "unknown input_kind");
  if (workspace_bytes < r2l_bwd_workspace_bytes(n_rays)) return fail("r2l_backward: %s", "workspace too small (see r2l_bwd_workspace_bytes)");
  if (misaligned(packed) || misaligned(workspace) || misaligned(fwd_saved) || misaligned(bwd_saved) || misaligned(zf) || misaligned(grads))
    return fail("r2l_backward: %s", "buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  r2l::ChainParams p;
  memset(&p, 0, sizeof(p));
  p.packed = static_cast<const uint8_t*>(packed);
  p.scratch = static_cast<float*>(workspace);
  p.saved = static_cast<uint8_t*>(bwd_saved);
  p.fwd_saved = static_cast<const uint8_t*>(fwd_saved);
  p.rgb_in = rgb;
  p.grad_rgb = grad_rgb;
  p.n_rays = n_rays;
  p.num_tiles = num_tiles(n_rays);
  p.input_kind = input_kind;
  set_debias(p);
  p.stats = g_stats;
  p.trace = g_trace;
  // When the chain grid (one CTA per tile) and the 90 weight-gradient CTAs fit on the GPU together, run them
  // concurrently: dw.cu starts on each layer as soon as every tile has stored that layer's dY operand.
  const int grid = fwd_grid(n_rays, r2l::kBwd);
  SideStream* side = (grid + 64 <= sm_count()) ? side_stream() : nullptr;
  int* ready = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + r2l_fwd_workspace_bytes(n_rays) - kReadyBytes);
  r2l::DwParams d;
  d.fwd_saved = p.fwd_saved;
  d.bwd_saved = p.saved;
  d.grads = grads;
  d.num_tiles = p.num_tiles;
  d.ready_target = p.num_tiles * (chain_form(r2l::kBwd, n_rays) == r2l::kFormHalf ? 2 : 1);   // store warps per tile
  d.input_kind = input_kind;
  d.accumulate = 0;
  d.ready = nullptr;
  d.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + r2l_fwd_workspace_bytes(n_rays));
  d.tickets = ready + 128;
  d.queue = ready + 240;                  // one claim counter per weight-gradient launch (<= kMaxGradChunks)
  d.times = g_trace ? g_trace + 148 * 5 * 96 : nullptr;   // the dW stamps follow the chain kernel's trace rows
  r2l::TailGradParams t;
  t.zf = zf;
  t.rgb = rgb;
  t.grad_rgb = grad_rgb;
  t.grads = grads;
  t.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + r2l_fwd_workspace_bytes(n_rays) + kDwPartialBytes);
  t.ticket = ready + 252;
  t.n_rays = n_rays;
  dw_schedule(d, side != nullptr);
  d.deterministic = g_deterministic;
  SideStream* ev_owner = n_chunks > 1 ? side_stream() : nullptr;
  if (n_chunks > 1 && !ev_owner) return fail("r2l_backward_chunked: %s", "no device state");
  // one weight-gradient launch per gradient chunk (units in release order), an event behind each but the last: the caller's
  // communication stream can reduce chunk i across ranks while the chain and the later launches still run.  The launches
  // leave `reserve_sms` SMs to that collective's kernel.
  auto launch_dw_chunks = [&](cudaStream_t s) -> int {
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int u0 = chunk_first_unit(n_chunks, split_layers, ch), u1 = chunk_first_unit(n_chunks, split_layers, ch + 1);
      r2l::DwParams dc = d;
      dc.item_lo = d.unit_first[u0];
      dc.item_hi = u1 < kDwUnits ? (int)d.unit_first[u1] : d.num_items;
      dc.queue = d.queue + ch;
      int g = dc.item_hi - dc.item_lo;
      // persistent CTAs beyond the SMs the chain leaves free simply start when SMs free up; the collective's kernel gets
      // its SMs at the launch boundaries through its stream's priority (trainer.py) and / or the SMs reserved here
      int cap = sm_count() - reserve_sms;
      if (cap < 8) cap = 8;
      dc.grid = g < cap ? g : cap;
      if (int rc = check_launch(r2l::launch_dw(dc, s), "r2l_backward(dw)")) return rc;
      if (ch + 1 < n_chunks)
        if (int rc = check(cudaEventRecord(ev_owner->chunk_done[ch], s), "r2l_backward(chunk event)")) return rc;
    }
    return 0;
  };
  // one small kernel instead of a memset: zeroes the flag words and derives the loss scale from max |grad_rgb|
  float* bwd_scale = reinterpret_cast<float*>(ready + 254);
  p.bwd_scale = bwd_scale;
  d.bwd_scale = bwd_scale;
  if (misaligned(grad_rgb)) return fail("r2l_backward: %s", "grad_rgb must be 16-byte aligned");
  if (int rc = check_launch(r2l::launch_bwd_prep(grad_rgb, n_rays * 3, ready, 254, bwd_scale, st), "r2l_backward(prep)")) return rc;
  // pieces of split units add into the buffer: head + body gradients start at 0 (zeroed on the stream dW runs on)
  const bool zero_grads = !d.deterministic && d.num_items > kDwUnits;
  if (side) {
    p.ready = ready;
    d.ready = ready;
    if (int rc = check(cudaEventRecord(side->fork, st), "r2l_backward(fork)")) return rc;
    if (int rc = check(cudaStreamWaitEvent(side->stream, side->fork, 0), "r2l_backward(fork wait)")) return rc;
    if (int rc = check_launch(launch_chain_any(r2l::kBwd, p, grid, st), "r2l_backward(chain)")) return rc;
    // the tail gradients need only forward results: first on the side stream, on SMs the chain leaves idle
    if (int rc = check_launch(r2l::launch_tail_grads(t, side->stream), "r2l_backward(tail)")) return rc;
    if (zero_grads)
      if (int rc = check(cudaMemsetAsync(grads, 0, (size_t)r2l::kOffTailW * sizeof(float), side->stream), "r2l_backward(zero grads)")) return rc;
    if (int rc = launch_dw_chunks(side->stream)) return rc;
    if (int rc = check(cudaEventRecord(side->join, side->stream), "r2l_backward(join)")) return rc;
    if (int rc = check(cudaStreamWaitEvent(st, side->join, 0), "r2l_backward(join wait)")) return rc;
  } else {
    if (int rc = check_launch(launch_chain_any(r2l::kBwd, p, grid, st), "r2l_backward(chain)")) return rc;
    if (zero_grads)
      if (int rc = check(cudaMemsetAsync(grads, 0, (size_t)r2l::kOffTailW * sizeof(float), st), "r2l_backward(zero grads)")) return rc;
    // the tail gradients belong to the first chunk (top of the buffer): before the weight-gradient launches here
    if (int rc = check_launch(r2l::launch_tail_grads(t, st), "r2l_backward(tail)")) return rc;
    if (int rc = launch_dw_chunks(st)) return rc;
  }
  return 0;
}
}  // namespace

int r2l_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int64_t n_rays, int n_samples,
                    int white_bkgd, float* rgb_map, float* disp_map, float* acc_map, float* weights, float* depth_map,
                    void* stream) {
  if (n_rays == 0) return 0;
  if (n_rays < 0 || n_samples <= 0) return fail("r2l_raw2outputs: %s", "bad sizes");
  if (!raw || !z_vals || !rays_d || !rgb_map || !disp_map || !acc_map || !weights || !depth_map)
    return fail("r2l_raw2outputs: %s", "null pointer");
  if (misaligned(raw)) return fail("r2l_raw2outputs: %s", "raw must be 16-byte aligned");
  return check_launch(r2l::launch_raw2outputs(raw, z_vals, rays_d, n_rays, n_samples, white_bkgd, rgb_map, disp_map, acc_map,
                                       weights, depth_map, (cudaStream_t)stream), "r2l_raw2outputs");
}

int r2l_sample_pdf_merge(const float* z_vals, const float* weights, const float* u, int64_t u_stride, int64_t n_rays,
                         int n_samples, int n_importance, float* z_samples, float* z_merged, void* stream) {
  if (n_rays == 0) return 0;
  if (n_rays < 0 || n_samples < 3 || n_importance < 1 || n_samples > 1024 || n_importance > 2048 || u_stride < 0)
    return fail("r2l_sample_pdf_merge: %s", "bad sizes");
  if (!z_vals || !weights || !u || !z_samples || !z_merged) return fail("r2l_sample_pdf_merge: %s", "null pointer");
  return check_launch(r2l::launch_sample_pdf_merge(z_vals, weights, u, u_stride, n_rays, n_samples, n_importance, z_samples, z_merged,
                                            nullptr, (cudaStream_t)stream), "r2l_sample_pdf_merge");
}

int r2l_sample_pdf(const float* bins, const float* weights, const float* u, int64_t u_stride, int64_t n_rays, int n_bins,
                   int n_importance, float* z_samples, void* stream) {
  if (n_rays == 0) return 0;
  if (n_rays < 0 || n_bins < 2 || n_importance < 1 || n_bins > 1023 || n_importance > 2048 || u_stride < 0)
    return fail("r2l_sample_pdf: %s", "bad sizes");
  if (!bins || !weights || !u || !z_samples) return fail("r2l_sample_pdf: %s", "null pointer");
  return check_launch(r2l::launch_sample_pdf_merge(nullptr, weights, u, u_stride, n_rays, n_bins + 1, n_importance, z_samples, nullptr,
                                            bins, (cudaStream_t)stream), "r2l_sample_pdf");
}

int r2l_positional_embed(const float* x, float* out, int64_t n, int dim, int n_freqs, int style, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || dim <= 0 || n_freqs <= 0 || n_freqs > 24 || (style != 0 && style != 1))
    return fail("r2l_positional_embed: %s", "bad arguments");
  if (!x || !out) return fail("r2l_positional_embed: %s", "null pointer");
  return check_launch(r2l::launch_embed(x, out, n, dim, n_freqs, style, (cudaStream_t)stream), "r2l_positional_embed");
}

size_t r2l_teacher_packed_bytes(void) { return (size_t)r2l::kTeacherPackedBytes; }

int r2l_teacher_pack_weights(const float* params, void* packed, void* stream) {
  if (!params || !packed) return fail("r2l_teacher_pack_weights: %s", "null pointer");
  return check_launch(r2l::launch_teacher_pack(params, packed, (cudaStream_t)stream), "r2l_teacher_pack_weights");
}

int r2l_teacher_forward(const float* pts, const float* viewdirs, const float* x_embedded, const void* packed, float* raw,
                        int64_t n_points, int64_t samples_per_ray, void* stream) {
  if (n_points == 0) return 0;
  if (n_points < 0) return fail("r2l_teacher_forward: %s", "negative n_points");
  if (!packed || !raw) return fail("r2l_teacher_forward: %s", "null pointer");
  if (!x_embedded && (!pts || !viewdirs || samples_per_ray <= 0))
    return fail("r2l_teacher_forward: %s", "needs pts + viewdirs + samples_per_ray, or x_embedded");
  if (misaligned(packed) || misaligned(raw)) return fail("r2l_teacher_forward: %s", "packed/raw must be 16-byte aligned");
  r2l::TeacherParams p;
  memset(&p, 0, sizeof(p));
  p.pts = pts;
  p.viewdirs = viewdirs;
  p.x_embedded = x_embedded;
  p.packed = static_cast<const uint8_t*>(packed);
  p.raw = raw;
  p.n_points = n_points;
  p.samples_per_ray = samples_per_ray > 0 ? samples_per_ray : 1;
  p.num_tiles = num_tiles(n_points);
  return check_launch(r2l::launch_teacher(p, plain_grid(n_points), (cudaStream_t)stream), "r2l_teacher_forward");
}

int r2l_teacher_forward_rays(const float* rays_o, const float* rays_d, const float* viewdirs, const float* z_vals, const void* packed,
                             float* raw, int64_t n_rays, int64_t samples_per_ray, void* stream) {
  if (n_rays == 0) return 0;
  if (n_rays < 0 || samples_per_ray <= 0) return fail("r2l_teacher_forward_rays: %s", "bad sizes");
  if (!rays_o || !rays_d || !viewdirs || !z_vals || !packed || !raw) return fail("r2l_teacher_forward_rays: %s", "null pointer");
  if (misaligned(packed) || misaligned(raw)) return fail("r2l_teacher_forward_rays: %s", "packed/raw must be 16-byte aligned");
  r2l::TeacherParams p;
  memset(&p, 0, sizeof(p));
  p.rays_o = rays_o;
  p.rays_d = rays_d;
  p.z_vals = z_vals;
  p.viewdirs = viewdirs;
  p.packed = static_cast<const uint8_t*>(packed);
  p.raw = raw;
  p.n_points = n_rays * samples_per_ray;
  p.samples_per_ray = samples_per_ray;
  p.num_tiles = num_tiles(p.n_points);
  return check_launch(r2l::launch_teacher(p, plain_grid(p.n_points), (cudaStream_t)stream), "r2l_teacher_forward_rays");
}

int r2l_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int64_t step, void* stream) {
  if (n == 0) return 0;
  if (!params || !grads || !exp_avg || !exp_avg_sq) return fail("r2l_adam_step: %s", "null pointer");
  if (n < 0 || step < 1) return fail("r2l_adam_step: %s", "bad n or step (steps count from 1)");
  if (misaligned(params) || misaligned(grads) || misaligned(exp_avg) || misaligned(exp_avg_sq))
    return fail("r2l_adam_step: %s", "buffers must be 16-byte aligned");
  // hyper-parameters arrive as doubles and are combined in double exactly as torch.optim.Adam does on the host
  const double bc1 = 1.0 - std::pow(beta1, (double)step);
  const double bc2 = 1.0 - std::pow(beta2, (double)step);
  return check_launch(r2l::launch_adam(params, grads, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                (float)eps, (float)(lr / bc1), (float)(1.0 / std::sqrt(bc2)), nullptr, (cudaStream_t)stream),
               "r2l_adam_step");
}

int r2l_adam_hyper(double lr, double beta1, double beta2, int64_t step, float* hyper_host) {
  if (!hyper_host || step < 1) return fail("r2l_adam_hyper: %s", "null pointer or step < 1");
  hyper_host[0] = (float)(lr / (1.0 - std::pow(beta1, (double)step)));
  hyper_host[1] = (float)(1.0 / std::sqrt(1.0 - std::pow(beta2, (double)step)));
  return 0;
}

int r2l_adam_schedule_dev(double lrate, double warmup_start_lr, double warmup_end_iter, double decay_rate, double decay_steps,
                          double beta1, double beta2, int64_t* counters, float* hyper, void* stream) {
  if (!counters || !hyper) return fail("r2l_adam_schedule_dev: %s", "null pointer");
  if (!(decay_steps > 0.0) || warmup_end_iter < 0.0) return fail("r2l_adam_schedule_dev: %s", "bad schedule");
  r2l::AdamSchedule sc{lrate, warmup_start_lr, warmup_end_iter, decay_rate, decay_steps, beta1, beta2};
  return check_launch(r2l::launch_adam_schedule(sc, reinterpret_cast<long long*>(counters), hyper, (cudaStream_t)stream), "r2l_adam_schedule_dev");
}

// ---- data-parallel state: one per process (= per GPU), see dp.cu ----
namespace {
struct DpState {
  bool active = false;
  int rank = 0, world = 1, device = -1;
  int64_t n = 0;
  void* local = nullptr;                 // one cudaMalloc block: [grads n | params n (padded to 256 B) | flags + state]
  void* peer[r2l::kDpMaxWorld] = {};     // IPC mappings of the other ranks' blocks (peer[rank] = local)
  size_t off_params = 0, off_flags = 0, bytes = 0;
} g_dp;
int g_dp_grid = 0, g_dp_variant = 0;   // debug overrides (r2l_debug_set_dp_grid)
size_t dp_round(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

size_t r2l_dp_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int r2l_dp_create(int rank, int world, int64_t n_params, void* handle_out) {
  if (g_dp.active) return fail("r2l_dp_create: %s", "already created in this process");
  if (world < 1 || world > r2l::kDpMaxWorld || rank < 0 || rank >= world || n_params <= 0 || !handle_out)
    return fail("r2l_dp_create: %s", "bad arguments");
  DpState d;
  d.rank = rank; d.world = world; d.n = n_params;
  if (int rc = check(cudaGetDevice(&d.device), "r2l_dp_create(device)")) return rc;
  d.off_params = dp_round((size_t)n_params * sizeof(float));
  d.off_flags = d.off_params + dp_round((size_t)n_params * sizeof(float));
  d.bytes = d.off_flags + 4096;          // flag words [0, 256): 8 launch slots x 32; state words from 512
  if (int rc = check(cudaMalloc(&d.local, d.bytes), "r2l_dp_create(cudaMalloc)")) return rc;
  if (int rc = check(cudaMemset(d.local, 0, d.bytes), "r2l_dp_create(memset)")) return rc;
  if (int rc = check(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle_out), d.local), "r2l_dp_create(cudaIpcGetMemHandle)")) return rc;
  d.peer[rank] = d.local;
  d.active = true;
  g_dp = d;
  return 0;
}

int r2l_dp_connect(const void* all_handles) {
  if (!g_dp.active || !all_handles) return fail("r2l_dp_connect: %s", "r2l_dp_create first / null pointer");
  const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int r = 0; r < g_dp.world; ++r) {
    if (r == g_dp.rank) continue;
    if (int rc = check(cudaIpcOpenMemHandle(&g_dp.peer[r], h[r], cudaIpcMemLazyEnablePeerAccess), "r2l_dp_connect(cudaIpcOpenMemHandle)")) return rc;
  }
  return 0;
}

void* r2l_dp_grads(void) { return g_dp.active ? g_dp.local : nullptr; }
void* r2l_dp_params(void) { return g_dp.active ? static_cast<uint8_t*>(g_dp.local) + g_dp.off_params : nullptr; }

namespace {
// slice of the float range [lo, hi) rank `g_dp.rank` owns: equal float4-aligned parts in rank order
void dp_slice(int64_t lo, int64_t hi, int64_t* slo, int64_t* shi) {
  const int64_t per = ((((hi - lo) + g_dp.world - 1) / g_dp.world) + 3) & ~(int64_t)3;
  int64_t a = lo + per * g_dp.rank, b = lo + per * (g_dp.rank + 1);
  *slo = a < hi ? a : hi;
  *shi = b < hi ? b : hi;
}
}  // namespace

int r2l_dp_slice(int64_t lo, int64_t hi, int64_t* slice_lo, int64_t* slice_hi) {
  if (!g_dp.active || !slice_lo || !slice_hi) return fail("r2l_dp_slice: %s", "r2l_dp_create first / null pointer");
  if (lo < 0 || hi > g_dp.n || lo > hi || (lo & 3)) return fail("r2l_dp_slice: %s", "bad range (lo must be a multiple of 4)");
  dp_slice(lo, hi, slice_lo, slice_hi);
  return 0;
}

int r2l_dp_adam_step_range(float* exp_avg, float* exp_avg_sq, double beta1, double beta2, double eps, const float* hyper,
                           int64_t lo, int64_t hi, int slot, int grid, void* stream) {
  if (!g_dp.active) return fail("r2l_dp_adam_step: %s", "r2l_dp_create / r2l_dp_connect first");
  if (!exp_avg || !exp_avg_sq || !hyper) return fail("r2l_dp_adam_step: %s", "null pointer");
  if (misaligned(exp_avg) || misaligned(exp_avg_sq)) return fail("r2l_dp_adam_step: %s", "moment buffers must be 16-byte aligned");
  if (lo < 0 || hi > g_dp.n || lo >= hi || (lo & 3) || slot < 0 || slot >= 8) return fail("r2l_dp_adam_step: %s", "bad range / slot");
  r2l::DpParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < g_dp.world; ++r) {
    if (!g_dp.peer[r]) return fail("r2l_dp_adam_step: %s", "a peer buffer is not connected");
    p.grads[r] = static_cast<float*>(g_dp.peer[r]);
    p.params[r] = reinterpret_cast<float*>(static_cast<uint8_t*>(g_dp.peer[r]) + g_dp.off_params);
    p.flags[r] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(g_dp.peer[r]) + g_dp.off_flags);
  }
  p.flag_base = 32 * slot;
  p.state = p.flags[g_dp.rank] + 512 + 2 * slot;
  p.exp_avg = exp_avg; p.exp_avg_sq = exp_avg_sq; p.hyper = hyper;
  dp_slice(lo, hi, &p.shard_lo, &p.shard_hi);
  p.w1 = (float)(1.0 - beta1); p.beta2 = (float)beta2; p.w2 = (float)(1.0 - beta2); p.eps = (float)eps;
  p.rank = g_dp.rank; p.world = g_dp.world;
  p.variant = g_dp_variant;
  if (grid <= 0) grid = 2 * sm_count();    // two CTAs of 256 threads per SM
  if (grid <= 0) grid = 148;
  if (g_dp_grid > 0) grid = g_dp_grid;
  return check_launch(r2l::launch_dp_adam(p, grid, (cudaStream_t)stream), "r2l_dp_adam_step");
}

int r2l_dp_adam_step(float* exp_avg, float* exp_avg_sq, double beta1, double beta2, double eps, const float* hyper, void* stream) {
  return r2l_dp_adam_step_range(exp_avg, exp_avg_sq, beta1, beta2, eps, hyper, 0, g_dp.active ? g_dp.n : 0, 0, 0, stream);
}

int r2l_debug_set_dp_grid(int grid, int variant) {
  g_dp_grid = grid;
  g_dp_variant = variant;
  return 0;
}

int r2l_dp_destroy(void) {
  if (!g_dp.active) return 0;
  cudaDeviceSynchronize();
  for (int r = 0; r < g_dp.world; ++r)
    if (r != g_dp.rank && g_dp.peer[r]) cudaIpcCloseMemHandle(g_dp.peer[r]);
  cudaFree(g_dp.local);
  g_dp = DpState();
  return 0;
}

int r2l_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double beta1, double beta2,
                      double eps, const float* hyper, void* stream) {
  if (n == 0) return 0;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !hyper) return fail("r2l_adam_step_dev: %s", "null pointer");
  if (n < 0) return fail("r2l_adam_step_dev: %s", "bad n");
  if (misaligned(params) || misaligned(grads) || misaligned(exp_avg) || misaligned(exp_avg_sq))
    return fail("r2l_adam_step_dev: %s", "buffers must be 16-byte aligned");
  return check_launch(r2l::launch_adam(params, grads, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                (float)eps, 0.f, 0.f, hyper, (cudaStream_t)stream), "r2l_adam_step_dev");
}

int r2l_read_ray_shards(const char* const* paths, int n_paths, float* dst_host, int64_t floats_per_shard, int n_threads) {
  if (n_paths == 0) return 0;
  if (!paths || !dst_host || n_paths < 0 || floats_per_shard <= 0) return fail("r2l_read_ray_shards: %s", "bad arguments");
  const std::string e = r2l::read_ray_shards(paths, n_paths, dst_host, floats_per_shard, n_threads);
  return e.empty() ? 0 : fail("r2l_read_ray_shards: %s", e.c_str());
}

size_t r2l_loss_scratch_bytes(void) { return 1024; }

int r2l_mse_loss_grad(const float* rgb, const float* target, int64_t n_rays, int target_stride, float grad_scale, float loss_scale,
                      float* grad_rgb, float* per_ray_err, float* loss, void* scratch, void* stream) {
  if (n_rays < 0 || target_stride < 3) return fail("r2l_mse_loss_grad: %s", "negative n_rays or target_stride < 3");
  if (!loss || !scratch || (n_rays > 0 && (!rgb || !target))) return fail("r2l_mse_loss_grad: %s", "null pointer");
  return check_launch(r2l::launch_mse_loss_grad(rgb, target, n_rays, target_stride, grad_scale, loss_scale, grad_rgb, per_ray_err, loss,
                                         static_cast<float*>(scratch), (cudaStream_t)stream), "r2l_mse_loss_grad");
}

int r2l_pool_draw(const float* pool_rows, const int32_t* pool_state, int n_out, uint64_t seed, const int64_t* counters,
                  float* dst_rows, int32_t* slots_out, void* stream) {
  if (n_out == 0) return 0;
  if (n_out < 0) return fail("r2l_pool_draw: %s", "negative n_out");
  if (!pool_rows || !pool_state || !counters || !dst_rows || !slots_out) return fail("r2l_pool_draw: %s", "null pointer");
  return check_launch(r2l::launch_pool_draw(pool_rows, pool_state, n_out, seed, reinterpret_cast<const long long*>(counters), dst_rows,
                                            slots_out, (cudaStream_t)stream), "r2l_pool_draw");
}

int64_t r2l_pool_slot_host(int64_t j, int64_t size, uint64_t seed, int64_t step) {
  if (size <= 0 || size > 0x7fffffff || j < 0 || j >= size) return -1;
  return (int64_t)r2l::pool_slot_host((uint32_t)j, (uint32_t)size, seed, (uint64_t)step);
}

int r2l_pool_update(const float* rays9, const float* per_ray_err, int64_t n_fresh, int n_hard_in, float* pool_rows, int32_t* pool_state,
                    const int32_t* slots_out, int32_t* picked, void* stream) {
  if (n_hard_in == 0) return 0;
  if (n_hard_in < 0 || n_fresh < n_hard_in || n_fresh > 0x7fffffff) return fail("r2l_pool_update: %s", "need 0 <= n_hard_in <= n_fresh < 2^31");
  if (!rays9 || !per_ray_err || !pool_rows || !pool_state) return fail("r2l_pool_update: %s", "null pointer");
  return check_launch(r2l::launch_pool_update(rays9, per_ray_err, (int)n_fresh, n_hard_in, pool_rows, pool_state, slots_out, picked,
                                              (cudaStream_t)stream), "r2l_pool_update");
}

int r2l_debug_set_stats(long long* stats) {
  g_stats = stats;
  return 0;
}

int r2l_set_deterministic(int on) {
  g_deterministic = on ? 1 : 0;
  return 0;
}

int r2l_debug_set_dw_schedule(int t1, int t2, int t3, int serial_pieces, int t4) {
  g_dw_sched[0] = t1; g_dw_sched[1] = t2; g_dw_sched[2] = t3; g_dw_sched[3] = serial_pieces; g_dw_sched[4] = t4;
  return 0;
}

int r2l_debug_set_accum_debias(float eps_body, float eps_head) {
  g_debias[0] = eps_body;
  g_debias[1] = eps_head;
  return 0;
}

long long r2l_debug_launch_count(int reset) {
  const long long n = g_launches;
  if (reset) g_launches = 0;
  return n;
}

int r2l_set_pair_mode(int mode) {
  g_form = (mode < 0 || mode > 2) ? -1 : mode;
  return 0;
}

int r2l_debug_set_trace(long long* trace) {
  g_trace = trace;
  return 0;
}

int r2l_debug_mma_rate(int form, int variant, int reps, int grid, long long* out_cycles, void* stream) {
  if (!out_cycles || reps < 1 || grid < 1 || form < 0 || form > 2 || (form > 0 && (grid & 1)))
    return fail("r2l_debug_mma_rate: %s", "bad arguments");
  return check_launch(r2l::launch_mma_rate(form, variant, reps, grid, out_cycles, (cudaStream_t)stream), "r2l_debug_mma_rate");
}

int r2l_selftest_layer(const float* A, const void* packed, int layer, float* C, void* stream) {
  if (!A || !packed || !C) return fail("r2l_selftest_layer: %s", "null pointer");
  if (layer < 0 || layer >= r2l::kBodyLayers) return fail("r2l_selftest_layer: %s", "layer out of range");
  const uint8_t* images = static_cast<const uint8_t*>(packed) + (int64_t)(r2l::kImgBody + 8 * layer) * r2l::kWImageBytes;
  return check_launch(r2l::launch_umma_selftest(A, images, C, (cudaStream_t)stream), "r2l_selftest_layer");
}

}  // extern "C"
