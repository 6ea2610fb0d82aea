// extern "C" boundary (include/r2l_b200.h). Argument checking, launch-geometry policy, error strings.
#include <cstdio>
#include <cstring>

#include "../../include/r2l_b200.h"
#include "kernels.cuh"

namespace {
thread_local char g_err[512] = "";
long long* g_stats = nullptr;  // debug cycle counters, see r2l_debug_set_stats

int fail(const char* fmt, const char* detail) {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return -1;
}
int check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -2;
}
int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}
int num_tiles(int64_t n_rays) { return (int)((n_rays + r2l::kTileM - 1) / r2l::kTileM); }
int fwd_grid(int64_t n_rays) {
  const int sms = sm_count();
  const int t = num_tiles(n_rays);
  return t < sms ? t : sms;
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
  const int t = num_tiles(n_rays);
  const int g = t < sms ? t : sms;
  return (size_t)(g > 0 ? g : 1) * r2l::kTileM * r2l::kWidth * sizeof(float);
}

int r2l_pack_weights(const float* params, void* packed, void* stream) {
  if (!params || !packed) return fail("r2l_pack_weights: %s", "null pointer");
  return check(r2l::launch_pack(params, packed, (cudaStream_t)stream), "r2l_pack_weights");
}

int r2l_forward(int input_kind, const float* in0, const float* in1, const float* t_rand, const float* z_lo,
                const float* z_diff, const void* packed, float* rgb, void* workspace, size_t workspace_bytes,
                int64_t n_rays, void* stream) {
  if (n_rays == 0) return 0;
  if (n_rays < 0) return fail("r2l_forward: %s", "negative n_rays");
  if (!in0 || !packed || !rgb || !workspace) return fail("r2l_forward: %s", "null pointer");
  if (input_kind < 0 || input_kind > 2) return fail("r2l_forward: %s", "unknown input_kind");
  if (input_kind == R2L_INPUT_RAYS && (!in1 || !z_lo)) return fail("r2l_forward: %s", "rays input needs in1 and z_lo");
  if (t_rand && (input_kind != R2L_INPUT_RAYS || !z_diff))
    return fail("r2l_forward: %s", "t_rand needs R2L_INPUT_RAYS and z_diff");
  if (workspace_bytes < r2l_fwd_workspace_bytes(n_rays)) return fail("r2l_forward: %s", "workspace too small");
  if (((uintptr_t)packed & 15) || ((uintptr_t)workspace & 15)) return fail("r2l_forward: %s", "packed/workspace must be 16-byte aligned");

  r2l::FwdParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = in0;
  p.in1 = in1;
  p.t_rand = t_rand;
  for (int i = 0; i < r2l::kSamples; ++i) {
    p.z_lo[i] = z_lo ? z_lo[i] : 0.f;
    p.z_diff[i] = z_diff ? z_diff[i] : 0.f;
  }
  p.packed = static_cast<const uint8_t*>(packed);
  p.rgb = rgb;
  p.h_scratch = static_cast<float*>(workspace);
  p.n_rays = n_rays;
  p.num_tiles = num_tiles(n_rays);
  p.input_kind = input_kind;
  p.stats = g_stats;
  return check(r2l::launch_fwd(p, fwd_grid(n_rays), (cudaStream_t)stream), "r2l_forward");
}

int r2l_debug_set_stats(long long* stats) {
  g_stats = stats;
  return 0;
}

int r2l_selftest_layer(const float* A, const void* packed, int layer, float* C, void* stream) {
  if (!A || !packed || !C) return fail("r2l_selftest_layer: %s", "null pointer");
  if (layer < 0 || layer >= r2l::kBodyLayers) return fail("r2l_selftest_layer: %s", "layer out of range");
  const uint8_t* images = static_cast<const uint8_t*>(packed) + (int64_t)(r2l::kImgBody + 8 * layer) * r2l::kWImageBytes;
  return check(r2l::launch_umma_selftest(A, images, C, (cudaStream_t)stream), "r2l_selftest_layer");
}

}  // extern "C"
