// Adam on the flat parameter buffer: one HBM-bound pass (reads p, g, m, v; writes p, m, v = 28 B/parameter).
// Reference: torch.optim.Adam(params, lr=args.lrate, betas=(0.9, 0.999)) at main.py:465 stepping 176 tensors
// (main.py:1406); same update rule, defaults amsgrad=False, weight_decay=0, maximize=False.  SURVEY.md row N3.
#include "kernels.cuh"

namespace r2l {

__global__ void __launch_bounds__(256) r2l_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                       float w1, float beta2, float w2, float eps, float step_size,
                                                       float inv_bc2_sqrt, const float* __restrict__ hyper) {   // w1 = 1 - beta1, w2 = 1 - beta2
  // hyper (optional, device): {step_size, inv_bc2_sqrt} of this step, so that a captured CUDA graph can be replayed with
  // the learning-rate schedule and the bias corrections of the current step (r2l_adam_step_dev)
  if (hyper != nullptr) {
    step_size = __ldg(hyper);
    inv_bc2_sqrt = __ldg(hyper + 1);
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ma[k] = ma[k] + (ga[k] - ma[k]) * w1;                 // exp_avg.lerp_(grad, 1 - beta1)
      va[k] = __fmaf_rn(w2 * ga[k], ga[k], va[k] * beta2);           // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
      const float denom = sqrtf(va[k]) * inv_bc2_sqrt + eps;           // (sqrt(v) / sqrt(bias_correction2)) + eps
      pa[k] = pa[k] - step_size * (ma[k] / denom);                     // param.addcdiv_(exp_avg, denom, -lr / bias_correction1)
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n is not a multiple of 4: 5,917,187 = 4 * 1,479,296 + 3)
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * w1;
    const float vi = __fmaf_rn(w2 * gi, gi, v[i] * beta2);
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - step_size * (mi / (sqrtf(vi) * inv_bc2_sqrt + eps));
  }
}

// Step scalars of the training iteration computed ON THE DEVICE from device-resident counters, so that a replayed CUDA
// graph (or a host that runs many iterations ahead of the GPU) applies exactly the learning rate and bias corrections of
// the iteration it executes.  counters[0] = global_step (the LR schedule's argument, main.py:1181-1195), counters[1] =
// Adam's step; both are incremented here, then
//   hyper[0] = lr / (1 - beta1^step)   hyper[1] = 1 / sqrt(1 - beta2^step)   hyper[2] = lr
// in double precision, as torch.optim.Adam and the reference's schedule compute them on the host.
__global__ void r2l_adam_schedule_kernel(AdamSchedule sc, long long* __restrict__ counters, float* __restrict__ hyper) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long gstep = ++counters[0];
  const long long astep = ++counters[1];
  double lr;
  if (sc.warmup_end > 0.0 && (double)gstep < sc.warmup_end)
    lr = (sc.lrate - sc.warmup_start_lr) / sc.warmup_end * (double)gstep + sc.warmup_start_lr;
  else
    lr = sc.lrate * pow(sc.decay_rate, ((double)gstep - sc.warmup_end) / sc.decay_steps);
  hyper[0] = (float)(lr / (1.0 - pow(sc.beta1, (double)astep)));
  hyper[1] = (float)(1.0 / sqrt(1.0 - pow(sc.beta2, (double)astep)));
  hyper[2] = (float)lr;
}

cudaError_t launch_adam_schedule(const AdamSchedule& sc, long long* counters, float* hyper, cudaStream_t stream) {
  r2l_adam_schedule_kernel<<<1, 32, 0, stream>>>(sc, counters, hyper);
  return cudaGetLastError();
}

cudaError_t launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float w1, float beta2, float w2, float eps,
                        float step_size, float inv_bc2_sqrt, const float* hyper, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  r2l_adam_kernel<<<sms * 8, 256, 0, stream>>>(p, g, m, v, n, w1, beta2, w2, eps, step_size, inv_bc2_sqrt, hyper);
  return cudaGetLastError();
}

}  // namespace r2l
