"""Rendering throughput (BASELINE configs 2 and 5): frames from poses (rays generated in-kernel, row N4) against the
reference's three-step idiom on the same library, for 400x400 and 800x800 frames and ray batches of 1k..64k."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import ops
from r2l_b200 import nerf_raybased as nb
dev = torch.device("cuda:0"); nb.device = dev
model = nb.NeRF_v3_2(nb.readme_args(), 1008, 3).to(dev)
with torch.no_grad(): model.flat.copy_(nb.init_flat_params(0).to(dev))
emb = nb.PositionalEmbedder(10)

def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

torch.manual_seed(0)
for H, P in ((400, 8), (800, 2)):
    focal = 555.5555155968841 * H / 400
    ps = nb.PointSampler(H, H, focal, 16, 2.0, 6.0)
    poses = torch.randn(P, 3, 4, device=dev) * 0.5; poses[:, :, 3] = torch.tensor([0., 0., 4.], device=dev)
    n = P * H * H
    with torch.no_grad():
        ms = timed(lambda: model.render_poses(poses, ps, focal))
        ms8 = timed(lambda: model.render_poses(poses, ps, focal, as_uint8=True))
        ms3 = timed(lambda: [model(emb(ps.sample_test(poses[k]))) for k in range(P)])
    print(f"{H}x{H} x {P} poses ({n} rays): render_poses fp32 {ms:.3f} ms = {n / ms / 1e3:.1f} M rays/s = {P / ms * 1e3:.0f} frames/s; "
          f"uint8 {ms8:.3f} ms; three-step idiom (sample_test -> embedder -> model) {ms3:.3f} ms = {n / ms3 / 1e3:.1f} M rays/s", flush=True)
packed = model.packed_weights(); z = ps.z_vals.tolist()
for n in (1024, 4096, 16384, 65536, 160000, 640000):
    o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); out = torch.empty(n, 3, device=dev)
    ms = timed(lambda: ops.forward(packed, rays_o=o, rays_d=d, z_vals=z, out=out), reps=10)
    print(f"forward {n} rays: {ms:.4f} ms = {n / ms / 1e3:.2f} M rays/s = {n * 11.789824e6 / ms / 1e9:.0f} TFLOP/s algorithmic", flush=True)
