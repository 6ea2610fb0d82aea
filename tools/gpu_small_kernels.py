"""CUDA-event timings of the small kernels of a train step at 4096 rays: pack, loss+grad, Adam (warm L2, back to back)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import ops
from r2l_b200.nerf_raybased import init_flat_params
dev = torch.device("cuda:0")
flat = init_flat_params(0).to(dev); packed = ops.pack_weights(flat)
n = 4096
rgb, tgt = torch.rand(n, 3, device=dev), torch.rand(n, 3, device=dev)
g, m, v = torch.randn_like(flat) * 1e-3, torch.zeros_like(flat), torch.zeros_like(flat)
h = torch.zeros(2).pin_memory(); ops.adam_hyper(5e-4, 0.9, 0.999, 1, h); hd = h.to(dev)

def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps

print(f"pack_weights        {timed(lambda: ops.pack_weights(flat, out=packed)):7.2f} us")
print(f"mse_loss_grad       {timed(lambda: ops.mse_loss_grad(rgb, tgt, 1e-4, 1e-4, want_per_ray=True)):7.2f} us")
print(f"adam_step_dev       {timed(lambda: ops.adam_step_dev(flat, g, m, v, 0.9, 0.999, 1e-8, hd)):7.2f} us")
print(f"adam + pack         {timed(lambda: (ops.adam_step_dev(flat, g, m, v, 0.9, 0.999, 1e-8, hd), ops.pack_weights(flat, out=packed))):7.2f} us")
