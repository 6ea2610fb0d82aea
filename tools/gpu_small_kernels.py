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

# ---- teacher-side kernels (config 4 sizes: a 32,768-ray chunk, 64 coarse / 192 fine samples) and the hard-ray pool ----
quick = len(sys.argv) > 1 and sys.argv[1] == "ncu"        # under ncu: each kernel ONCE (the printed times mean nothing then)
reps = 1 if quick else 20
if quick:
    def timed(fn, reps=1):
        fn()
        torch.cuda.synchronize()
        return 1.0
from r2l_b200 import nerf_raybased as nb
nb.device = dev
torch.manual_seed(0)
HBM = 6552.0
for S in (64, 192):
    N = 32768
    raw = torch.randn(N, S, 4, device=dev); zv = torch.sort(torch.rand(N, S, device=dev) * 4 + 2, dim=-1).values; rd = torch.randn(N, 3, device=dev)
    us = timed(lambda: ops.raw2outputs(raw, zv, rd, True), reps)
    nbytes = N * (4 * (5 * S + 3) + 4 * (6 + S))
    print(f"raw2outputs {N} x {S}: {us:7.2f} us = {nbytes / us / 1e3:6.0f} GB/s algorithmic = {nbytes / us / 1e3 / HBM:.2f} of the HBM peak (incl. output allocation)")
N, S, M = 32768, 64, 128
zv = torch.sort(torch.rand(N, S, device=dev) * 4 + 2, dim=-1).values; w = torch.rand(N, S, device=dev)
us = timed(lambda: ops.sample_pdf_merge(zv, w, M, None), reps)
nbytes = N * (4 * 2 * S + 4 * (2 * M + S))
print(f"sample_pdf_merge {N} x {S} -> {M}: {us:7.2f} us = {nbytes / us / 1e3:6.0f} GB/s algorithmic")
teacher = nb.NeRF(8, 256, 63, 27, 4, [4], True).to(dev)
N, S = (2048, 192) if quick else (32768, 192)
o = torch.randn(N, 3, device=dev) * 0.1 + torch.tensor([0., 0., 4.], device=dev); d = torch.randn(N, 3, device=dev) * 0.3
zv = torch.sort(torch.rand(N, S, device=dev) * 4 + 2, dim=-1).values; vd = d / d.norm(dim=-1, keepdim=True)
with torch.no_grad():
    pts = (o[:, None, :] + d[:, None, :] * zv[..., None]).contiguous()
    us_p = timed(lambda: teacher.query(pts, vd), 2 if quick else 5)
    us_r = timed(lambda: teacher.query_rays(o, d, zv, vd), 2 if quick else 5)
    us_m = timed(lambda: (o[:, None, :] + d[:, None, :] * zv[..., None]), reps)
print(f"teacher query {N} x {S}: on points {us_p / 1e3:.3f} ms (+ {us_m / 1e3:.3f} ms to build them in torch), on rays {us_r / 1e3:.3f} ms = "
      f"{N * S / us_r:.1f} M points/s = {N * S * 2 * 593408 / us_r / 1e6:.0f} TFLOP/s algorithmic")
for nfresh, k in ((4096, 819), (81920, 16384)):
    rays9 = torch.rand(nfresh + k, 9, device=dev); err = torch.rand(nfresh + k, device=dev)
    pool = torch.zeros(20 * nfresh, 9, device=dev); state = torch.tensor([10 * nfresh], dtype=torch.int32, device=dev)
    slots = torch.zeros(k, dtype=torch.int32, device=dev); counters = torch.tensor([5, 5], dtype=torch.int64, device=dev)
    us_d = timed(lambda: ops.pool_draw(pool, state, k, 0, counters, rays9[nfresh:], slots), reps)
    us_u = timed(lambda: ops.pool_update(rays9, err, nfresh, k, pool, state, slots), reps)
    print(f"hard-ray pool, {nfresh} fresh rays, {k} in / out: draw {us_d:6.2f} us, update (select + scatter) {us_u:6.2f} us")
