"""Calibration of the accumulation debias of the half form (include/r2l_b200.h: r2l_debug_set_accum_debias).

The tensor core rounds its fp32 accumulator toward zero after every tcgen05.mma, so a GEMM issued as 48 (body) or 192 (head)
instructions comes out short by a few ulp on average.  This tool sweeps the two correction factors and reports the forward
error against the fp64 oracle (max and rms relative RGB error on lego-pose and stress rays) and the flat gradient error
against the fp64 autograd at 200 / 1000 rays, so that the minimum can be read off.  Usage: python tools/gpu_accum_calibrate.py"""
import importlib.util, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import r2l_oracle as orc
from oracle.torch_reference import RefR2L, embed, sample
from r2l_b200 import ops, _lib
from r2l_b200.nerf_raybased import init_flat_params
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
dev = torch.device("cuda:0")
L = _lib.lib()
flat = init_flat_params(0); packed = ops.pack_weights(flat.to(dev))
z = orc.sampler_z_vals(2.0, 6.0); zt = torch.from_numpy(z).to(dev)
U = 2.0 ** -24

def batches():
    out = {}
    ro, rd, tg = bench.synthetic_rays(1024, 0)
    out["lego"] = (torch.from_numpy(ro), torch.from_numpy(rd), torch.from_numpy(tg))
    torch.manual_seed(1000)
    out["stress"] = (torch.randn(1000, 3) * 0.5, torch.randn(1000, 3), torch.rand(1000, 3))
    torch.manual_seed(200)
    out["stress200"] = (torch.randn(200, 3) * 0.5, torch.randn(200, 3), torch.rand(200, 3))
    return out

B = batches()
truth = {}
for name, (o, d, t) in B.items():
    ref = RefR2L().load_flat(flat).double().to(dev)
    rgb64 = ref(embed(sample(o.to(dev), d.to(dev), zt)).double())
    ((rgb64 - t.to(dev).double()) ** 2).mean().backward()
    truth[name] = (rgb64.detach(), ref.flat_grads())

def measure(eb, eh):
    L.r2l_debug_set_accum_debias(eb, eh)
    row = []
    for name, (o, d, t) in B.items():
        o, d, t = o.to(dev), d.to(dev), t.to(dev)
        rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z.tolist())
        n = o.shape[0]
        g = ops.backward(packed, ctx, (2.0 / (3 * n)) * (rgb - t)).double()
        rgb64, g64 = truth[name]
        e = ((rgb.double() - rgb64) / rgb64)
        row.append((float(e.abs().max()), float(e.pow(2).mean().sqrt()), float(e.mean()), float((g - g64).norm() / g64.norm())))
    return row

print("eps in units of 2^-24; per batch: fwd max rel | fwd rms rel | fwd mean signed rel | flat gradient rel error")
print("batches:", list(B))
for eh in (0, 16, 32, 48, 64, 96):
    r = measure(0.0, eh * U)
    print(f"body  0 head {eh:3d}: " + "  ".join(f"{a:.2e}|{b:.2e}|{c:+.1e}|{g:.2e}" for a, b, c, g in r), flush=True)
for eb in (0, 2, 4, 6, 8, 10, 12, 16, 20, 24, 32):
    for eh in (0, 32, 64):
        r = measure(eb * U, eh * U)
        print(f"body {eb:2d} head {eh:3d}: " + "  ".join(f"{a:.2e}|{b:.2e}|{c:+.1e}|{g:.2e}" for a, b, c, g in r), flush=True)
L.r2l_debug_set_accum_debias(0.0, 0.0)
