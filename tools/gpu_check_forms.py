"""The three launch forms of the chain kernels (0 single, 1 pair, 2 half; chain.cu): parity against the golden vectors /
each other, and speed.  Usage: python tools/gpu_check_forms.py [forms, default 0,1,2]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2l_oracle as orc
from r2l_b200 import ops, _lib
from r2l_b200.nerf_raybased import init_flat_params
dev = torch.device("cuda:0")
forms = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2").split(",")]
flat = init_flat_params(0); packed = ops.pack_weights(flat.to(dev))
g = dict(np.load(os.path.join(ROOT, "tests/golden/r2l_seed0.npz")))
L = _lib.lib()
z = orc.sampler_z_vals(2.0, 6.0).tolist()
def rel(a, b): return float(np.max(np.abs(a - b) / np.abs(b)))
ro, rd = torch.from_numpy(g["rays_o"]).to(dev), torch.from_numpy(g["rays_d"]).to(dev)
for f in forms:
    L.r2l_set_pair_mode(f)
    rgb = ops.forward(packed, rays_o=ro, rays_d=rd, z_vals=g["z_vals"].tolist()); torch.cuda.synchronize()
    print(f"form {f} golden fwd (200 rays): max rel err {rel(rgb.cpu().numpy(), g['rgb']):.3e}", flush=True)
for n in (100, 129, 1000, 4096, 20001):
    torch.manual_seed(n); o = (torch.randn(n, 3) * 0.5).to(dev); d = torch.randn(n, 3).to(dev)
    L.r2l_set_pair_mode(0); a = ops.forward(packed, rays_o=o, rays_d=d, z_vals=z)
    for f in forms:
        L.r2l_set_pair_mode(f); b = ops.forward(packed, rays_o=o, rays_d=d, z_vals=z); torch.cuda.synchronize()
        print(f"N={n}: form {f} vs single max abs diff {float((a - b).abs().max()):.3e}  finite={bool(torch.isfinite(b).all())}", flush=True)
# training path
for n in (1100, 4096):
    torch.manual_seed(1); o = (torch.randn(n, 3) * 0.5).to(dev); d = torch.randn(n, 3).to(dev); t = torch.rand(n, 3).to(dev)
    grads = {}
    for f in sorted(set([0] + forms)):
        L.r2l_set_pair_mode(f)
        rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z)
        grads[f] = ops.backward(packed, ctx, (2.0 / (3 * n)) * (rgb - t)).clone(); torch.cuda.synchronize()
        print(f"train N={n}: grads form {f} vs single rel diff {float((grads[0] - grads[f]).norm() / grads[0].norm()):.3e}", flush=True)
for f in forms:
    L.r2l_set_pair_mode(f)
    for n in (4096, 9472, 18944, 160000):
        o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); out = torch.empty(n, 3, device=dev)
        for _ in range(3): ops.forward(packed, rays_o=o, rays_d=d, z_vals=z, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.forward(packed, rays_o=o, rays_d=d, z_vals=z, out=out)
        e1.record(); torch.cuda.synchronize()
        print(f"form {f} fwd N={n}: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
    n = 4096
    o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); t = torch.rand(n, 3, device=dev); gr = torch.empty(ops.NUM_PARAMS, device=dev)
    def fwd():
        return ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z)
    def step():
        rgb, ctx = fwd(); ops.backward(packed, ctx, (2.0 / (3 * n)) * (rgb - t), gr)
    for name, fn in (("fwd_train", fwd), ("fwd+bwd", step)):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"form {f} {name} N=4096: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
L.r2l_set_pair_mode(-1)
