"""GPU diagnostic for the training path: fused forward_train + backward vs the fp64 oracle / torch autograd."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2l_oracle as orc
from oracle.torch_reference import RefR2L, embed, sample
from r2l_b200 import ops
from r2l_b200.nerf_raybased import init_flat_params, state_dict_layout

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
flat = init_flat_params(0); flat_np = flat.numpy(); flat_d = flat.to(dev)
packed = ops.pack_weights(flat_d)
g = dict(np.load(os.path.join(ROOT, "tests/golden/r2l_seed0.npz")))
layout = state_dict_layout()

def report(tag, ours, ref64, ref32_err=None):
    ours = ours.astype(np.float64)
    print(f"{tag}: flat Frobenius rel err vs fp64 = {np.linalg.norm(ours - ref64) / np.linalg.norm(ref64):.3e}", flush=True)
    worst = (0, "")
    groups = {"head.0.weight": [], "head.0.bias": [], "body.w": [], "body.b": [], "tail.0.weight": [], "tail.0.bias": []}
    for name, shape, off in layout:
        n = int(np.prod(shape))
        e = np.linalg.norm(ours[off:off + n] - ref64[off:off + n]) / max(np.linalg.norm(ref64[off:off + n]), 1e-300)
        key = name if name in groups else ("body.w" if name.endswith("weight") else "body.b")
        groups[key].append(e)
        if e > worst[0]: worst = (e, name)
    print("   per-tensor rel err max by group:", {k: f"{max(v):.2e}" for k, v in groups.items()}, "worst:", worst[1], flush=True)

# ---- golden 200 rays: exact fp64 gradients from the oracle ----
ro, rd = torch.from_numpy(g["rays_o"]).to(dev), torch.from_numpy(g["rays_d"]).to(dev)
tgt = torch.from_numpy(g["target"]).to(dev)
rgb, ctx = ops.forward_train(packed, rays_o=ro, rays_d=rd, z_vals=g["z_vals"].tolist())
torch.cuda.synchronize()
print("train fwd rgb max rel err vs reference:", float(np.max(np.abs(rgb.cpu().numpy() - g["rgb"]) / g["rgb"])), flush=True)
zf_ref = None
grad_rgb = (2.0 / (3 * 200)) * (rgb - tgt)
grads = ops.backward(packed, ctx, grad_rgb)
torch.cuda.synchronize()
print("backward ran", flush=True)
loss64, g64, _, _ = orc.r2l_loss_and_grads(flat_np.astype(np.float64), g["x_embed"].astype(np.float64), g["target"].astype(np.float64))
report("golden N=200", grads.cpu().numpy(), g64)
print("   reference fp32 autograd vs fp64 (same metric):", float(g["grad_f32_vs_f64_rel"]), flush=True)
gn = grads.cpu().numpy()
for name in ("tail.0.weight", "tail.0.bias", "head.0.bias", "body.0.body.0.bias", "body.42.body.2.bias"):
    ref = g["g64_" + name]; off = [o for n_, s, o in layout if n_ == name][0]
    print(f"   {name}: rel err {np.linalg.norm(gn[off:off+ref.size]-ref)/np.linalg.norm(ref):.3e}", flush=True)

# ---- N = 4096 and 1000 (ragged): fp64 torch autograd on the GPU as truth ----
ref64 = RefR2L().load_flat(flat).double().to(dev)
zt = torch.from_numpy(orc.sampler_z_vals(2.0, 6.0)).to(dev)
for n in (1000, 4096):
    torch.manual_seed(n)
    o = (torch.randn(n, 3) * 0.5).to(dev); d = torch.randn(n, 3).to(dev); t = torch.rand(n, 3).to(dev)
    rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=zt.tolist())
    grads = ops.backward(packed, ctx, (2.0 / (3 * n)) * (rgb - t))
    torch.cuda.synchronize()
    ref64.zero_grad()
    x64 = embed(sample(o, d, zt)).double()
    ((ref64(x64) - t.double()) ** 2).mean().backward()
    report(f"N={n}", grads.cpu().numpy(), ref64.flat_grads().cpu().numpy())
    ref32 = RefR2L().load_flat(flat).to(dev)
    ((ref32(embed(sample(o, d, zt))) - t) ** 2).mean().backward()
    report(f"N={n} [stock torch fp32 autograd, for scale]", ref32.flat_grads().cpu().numpy(), ref64.flat_grads().cpu().numpy())

# ---- timing ----
for n in (4096, 18944, 98304):
    o = (torch.randn(n, 3, device=dev) * 0.5); d = torch.randn(n, 3, device=dev); t = torch.rand(n, 3, device=dev)
    grads = torch.empty(ops.NUM_PARAMS, device=dev)
    def step():
        rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=zt.tolist())
        ops.backward(packed, ctx, (2.0 / (3 * n)) * (rgb - t), grads)
    for _ in range(2): step()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    evs[0].record(); rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=zt.tolist()); evs[1].record()
    gr = (2.0 / (3 * n)) * (rgb - t)
    evs[2].record(); ops.backward(packed, ctx, gr, grads); evs[3].record(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"fwd+bwd N={n}: {ms:.3f} ms  ({n / ms / 1e3:.2f} Mrays/s)   fwd_train {evs[0].elapsed_time(evs[1]):.3f} ms  backward(all 3 kernels) {evs[2].elapsed_time(evs[3]):.3f} ms", flush=True)
