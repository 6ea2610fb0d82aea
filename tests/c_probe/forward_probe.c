/* A plain C99 host program that drives the hot path through include/r2l_b200.h alone - what a non-Python binding of the
 * library does: cudaMalloc / cudaMemcpy from the CUDA runtime's C API, r2l_pack_weights, r2l_forward, and (second part)
 * r2l_forward_train -> r2l_mse_loss_grad -> r2l_backward.  No torch, no Python, no C++.
 *
 *   forward_probe params.f32 rays_o.f32 rays_d.f32 target.f32 z_vals.f32 n_rays out_rgb.f32 out_grads.f32
 *
 * Inputs are raw little-endian float32 files; z_vals.f32 holds PointSampler's 16 depths (model/nerf_raybased.py:87-88).
 * Exit code 0 on success; every failure prints r2l_last_error(). */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/r2l_b200.h"

static float* read_f32(const char* path, size_t n) {
  float* p = (float*)malloc(n * sizeof(float));
  FILE* f = fopen(path, "rb");
  if (!f || fread(p, sizeof(float), n, f) != n) { fprintf(stderr, "cannot read %zu floats from %s\n", n, path); exit(2); }
  fclose(f);
  return p;
}
static void write_f32(const char* path, const float* p, size_t n) {
  FILE* f = fopen(path, "wb");
  if (!f || fwrite(p, sizeof(float), n, f) != n) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
  fclose(f);
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 3; } } while (0)
#define R2L(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #x, r_, r2l_last_error()); return 4; } } while (0)

static void* dev_alloc(size_t bytes) {
  void* p = NULL;
  if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { fprintf(stderr, "cudaMalloc(%zu) failed\n", bytes); exit(3); }
  return p;
}

int main(int argc, char** argv) {
  if (argc != 9) { fprintf(stderr, "usage: see the header comment\n"); return 1; }
  const long n = atol(argv[6]);
  const float* z_vals = read_f32(argv[5], 16);
  float* h_params = read_f32(argv[1], R2L_NUM_PARAMS);
  float* h_o = read_f32(argv[2], 3 * (size_t)n);
  float* h_d = read_f32(argv[3], 3 * (size_t)n);
  float* h_t = read_f32(argv[4], 3 * (size_t)n);

  cudaStream_t stream;
  CU(cudaStreamCreate(&stream));
  float* d_params = (float*)dev_alloc(R2L_NUM_PARAMS * sizeof(float));
  float* d_o = (float*)dev_alloc(3 * n * sizeof(float));
  float* d_d = (float*)dev_alloc(3 * n * sizeof(float));
  float* d_t = (float*)dev_alloc(3 * n * sizeof(float));
  float* d_rgb = (float*)dev_alloc(3 * n * sizeof(float));
  void* d_packed = dev_alloc(r2l_packed_bytes());
  const size_t ws_bytes = r2l_bwd_workspace_bytes(n);
  void* d_ws = dev_alloc(ws_bytes);
  CU(cudaMemcpyAsync(d_params, h_params, R2L_NUM_PARAMS * sizeof(float), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(d_o, h_o, 3 * n * sizeof(float), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(d_d, h_d, 3 * n * sizeof(float), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(d_t, h_t, 3 * n * sizeof(float), cudaMemcpyHostToDevice, stream));

  /* inference: NeRF_v3_2.forward(PositionalEmbedder(PointSampler.sample_train(rays_o, rays_d, perturb = 0))) */
  R2L(r2l_pack_weights(d_params, d_packed, stream));
  R2L(r2l_forward(R2L_INPUT_RAYS, d_o, d_d, NULL, z_vals, NULL, d_packed, d_rgb, d_ws, ws_bytes, n, stream));
  float* h_rgb = (float*)malloc(3 * n * sizeof(float));
  CU(cudaMemcpyAsync(h_rgb, d_rgb, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  write_f32(argv[7], h_rgb, 3 * (size_t)n);

  /* training: forward that keeps the operands, img2mse and its gradient, backward -> flat gradient in state_dict order */
  float* d_zf = (float*)dev_alloc(256 * n * sizeof(float));
  void* d_fwd_saved = dev_alloc(r2l_train_fwd_saved_bytes(n));
  void* d_bwd_saved = dev_alloc(r2l_train_bwd_saved_bytes(n));
  float* d_grad_rgb = (float*)dev_alloc(3 * n * sizeof(float));
  float* d_grads = (float*)dev_alloc(R2L_NUM_PARAMS * sizeof(float));
  float* d_loss = (float*)dev_alloc(sizeof(float));
  void* d_scratch = dev_alloc(r2l_loss_scratch_bytes());
  CU(cudaMemsetAsync(d_scratch, 0, r2l_loss_scratch_bytes(), stream));
  R2L(r2l_forward_train(R2L_INPUT_RAYS, d_o, d_d, NULL, z_vals, NULL, d_packed, d_rgb, d_zf, d_fwd_saved, d_ws, ws_bytes, n, stream));
  R2L(r2l_mse_loss_grad(d_rgb, d_t, n, 3, 2.0f / (3.0f * (float)n), 1.0f / (3.0f * (float)n), d_grad_rgb, NULL, d_loss, d_scratch, stream));
  R2L(r2l_backward(R2L_INPUT_RAYS, d_packed, d_rgb, d_grad_rgb, d_zf, d_fwd_saved, d_bwd_saved, d_grads, d_ws, ws_bytes, n, stream));
  float* h_grads = (float*)malloc(R2L_NUM_PARAMS * sizeof(float));
  float h_loss = 0.f;
  CU(cudaMemcpyAsync(h_grads, d_grads, R2L_NUM_PARAMS * sizeof(float), cudaMemcpyDeviceToHost, stream));
  CU(cudaMemcpyAsync(&h_loss, d_loss, sizeof(float), cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  write_f32(argv[8], h_grads, R2L_NUM_PARAMS);
  printf("abi %d rays %ld loss %.9g launches %lld\n", r2l_abi_version(), n, (double)h_loss, r2l_debug_launch_count(0));
  return 0;
}
