"""CTA-pair (cta_group::2) chain kernels vs the single-CTA ones: parity and speed."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import r2l_oracle as orc
from r2l_b200 import ops, _lib
from r2l_b200.nerf_raybased import init_flat_params
dev = torch.device("cuda:0")
flat = init_flat_params(0); packed = ops.pack_weights(flat.to(dev))
g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/r2l_seed0.npz")))
L = _lib.lib()
z = orc.sampler_z_vals(2.0, 6.0).tolist()
def rel(a, b): return float(np.max(np.abs(a - b) / np.abs(b)))
ro, rd = torch.from_numpy(g["rays_o"]).to(dev), torch.from_numpy(g["rays_d"]).to(dev)
L.r2l_set_pair_mode(1)
rgb = ops.forward(packed, rays_o=ro, rays_d=rd, z_vals=g["z_vals"].tolist()); torch.cuda.synchronize()
print("pair golden fwd (200 rays, 2 tiles): max rel err", rel(rgb.cpu().numpy(), g["rgb"]), flush=True)
for n in (100, 129, 1000, 4096, 20000):
    torch.manual_seed(n); o = (torch.randn(n, 3) * 0.5).to(dev); d = torch.randn(n, 3).to(dev)
    L.r2l_set_pair_mode(0); a = ops.forward(packed, rays_o=o, rays_d=d, z_vals=z)
    L.r2l_set_pair_mode(1); b = ops.forward(packed, rays_o=o, rays_d=d, z_vals=z); torch.cuda.synchronize()
    print(f"N={n}: pair vs single max abs diff {float((a - b).abs().max()):.3e}  finite={bool(torch.isfinite(b).all())}", flush=True)
# training path
n = 4096
torch.manual_seed(1); o = (torch.randn(n, 3) * 0.5).to(dev); d = torch.randn(n, 3).to(dev); t = torch.rand(n, 3).to(dev)
grads = {}
for mode in (0, 1):
    L.r2l_set_pair_mode(mode)
    rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z)
    grads[mode] = ops.backward(packed, ctx, (2.0 / (3 * n)) * (rgb - t)).clone(); torch.cuda.synchronize()
print("train N=4096: grads pair vs single rel diff", float((grads[0] - grads[1]).norm() / grads[0].norm()), flush=True)
for mode in (0, 1):
    L.r2l_set_pair_mode(mode)
    for n in (4096, 18944, 160000):
        o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); out = torch.empty(n, 3, device=dev)
        for _ in range(3): ops.forward(packed, rays_o=o, rays_d=d, z_vals=z, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.forward(packed, rays_o=o, rays_d=d, z_vals=z, out=out)
        e1.record(); torch.cuda.synchronize()
        print(f"pair={mode} fwd N={n}: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
    n = 4096
    o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); t = torch.rand(n, 3, device=dev); gr = torch.empty(ops.NUM_PARAMS, device=dev)
    def step():
        rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z); ops.backward(packed, ctx, (2.0 / (3 * n)) * (rgb - t), gr)
    for _ in range(3): step()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    print(f"pair={mode} fwd+bwd N=4096: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
