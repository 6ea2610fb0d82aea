"""GPU diagnostic: single-layer tcgen05 self test, forward parity vs the oracle, quick timing. Run under gpurun."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2l_oracle as orc
from r2l_b200 import ops
from r2l_b200.nerf_raybased import init_flat_params

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)
flat = init_flat_params(0)
flat_np = flat.numpy()
flat_d = flat.to(dev)
packed = ops.pack_weights(flat_d)
torch.cuda.synchronize()
print("pack ok", packed.numel(), flush=True)

# ---- pack check: decode image of body layer 3 chunk 1 and compare hi+lo with W ----
P = packed.cpu().numpy()
def decode_image(img_bytes):  # [256 n][64 k] from a swizzled 16-bit plane
    raw = np.frombuffer(img_bytes, dtype=np.uint16).reshape(256, 8, 8)  # row, 16B-unit position, elem
    out = np.zeros((256, 64), np.float32)
    for n in range(256):
        for pos in range(8):
            j = pos ^ (n & 7)
            out[n, 8 * j:8 * j + 8] = (raw[n, pos].astype(np.uint32) << 16).view(np.float32)
    return out
l, c = 3, 1
base = (64 + 8 * l + 2 * c) * 32768
hi = decode_image(P[base:base + 32768].tobytes()); lo = decode_image(P[base + 32768:base + 65536].tobytes())
W = orc.unflatten_params(flat_np)["body"][l][0]
print("pack image err", np.abs(hi + lo - W[:, 64 * c:64 * c + 64]).max(), "hi-only err", np.abs(hi - W[:, 64 * c:64 * c + 64]).max(), flush=True)

# ---- single-layer UMMA self test ----
torch.manual_seed(5)
A = torch.randn(128, 256)
for layer in (0, 7):
    C = ops.selftest_layer(A.to(dev), packed, layer)
    torch.cuda.synchronize()
    Wl = orc.unflatten_params(flat_np)["body"][layer][0].astype(np.float64)
    ref = A.numpy().astype(np.float64) @ Wl.T
    err = np.abs(C.cpu().numpy() - ref).max() / np.abs(ref).max()
    print(f"selftest layer {layer}: max err / max |ref| = {err:.3e}", flush=True)

# ---- forward parity ----
g = dict(np.load(os.path.join(ROOT, "tests/golden/r2l_seed0.npz")))
def relerr(a, b): return float(np.max(np.abs(a - b) / np.abs(b)))
ro, rd = torch.from_numpy(g["rays_o"]).to(dev), torch.from_numpy(g["rays_d"]).to(dev)
rgb = ops.forward(packed, rays_o=ro, rays_d=rd, z_vals=g["z_vals"].tolist()); torch.cuda.synchronize()
print("golden rays   : max rel err vs reference rgb", relerr(rgb.cpu().numpy(), g["rgb"]), flush=True)
rgb = ops.forward(packed, pts=torch.from_numpy(g["pts"]).to(dev)); torch.cuda.synchronize()
print("golden pts    : max rel err", relerr(rgb.cpu().numpy(), g["rgb"]), flush=True)
rgb = ops.forward(packed, x=torch.from_numpy(g["x_embed"]).to(dev)); torch.cuda.synchronize()
print("golden x      : max rel err", relerr(rgb.cpu().numpy(), g["rgb"]), flush=True)
lo_, df_ = orc.jitter_bounds(g["z_vals"])
rgb = ops.forward(packed, rays_o=ro, rays_d=rd, t_rand=torch.from_numpy(g["t_rand"]).to(dev), z_lower=lo_.tolist(), z_diff=df_.tolist()); torch.cuda.synchronize()
print("golden jitter : max rel err", relerr(rgb.cpu().numpy(), g["rgb_jit"]), flush=True)

for n in (1, 127, 128, 129, 4096, 20000):
    torch.manual_seed(1)
    o = torch.randn(n, 3) * 0.5; d = torch.randn(n, 3)
    z = orc.sampler_z_vals(2.0, 6.0)
    rgb = ops.forward(packed, rays_o=o.to(dev), rays_d=d.to(dev), z_vals=z.tolist()); torch.cuda.synchronize()
    m = min(n, 512)
    x = orc.positional_embed(orc.sample_train(o.numpy()[:m], d.numpy()[:m], z, None))
    ref = orc.r2l_forward(flat_np, x)
    print(f"stress N={n}: max rel err (first {m}) {relerr(rgb.cpu().numpy()[:m], ref):.3e}  finite={bool(torch.isfinite(rgb).all())}", flush=True)

# ---- timing ----
import ctypes
from r2l_b200 import _lib
stats = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
for n in (4096, 18944, 160000):
    o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev)
    z = orc.sampler_z_vals(2.0, 6.0).tolist()
    out = torch.empty(n, 3, device=dev)
    for _ in range(3): ops.forward(packed, rays_o=o, rays_d=d, z_vals=z, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps): ops.forward(packed, rays_o=o, rays_d=d, z_vals=z, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"fwd N={n}: {ms:.4f} ms  {n / ms / 1e3:.2f} Mrays/s  {n * 11.789824e6 / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)
    stats.zero_(); _lib.lib().r2l_debug_set_stats(ctypes.c_void_p(stats.data_ptr()))
    ops.forward(packed, rays_o=o, rays_d=d, z_vals=z, out=out); torch.cuda.synchronize()
    _lib.lib().r2l_debug_set_stats(None)
    st = stats.view(148, 8).cpu().numpy().astype(np.float64)
    act = st[:, 4] > 0
    tiles = np.ceil(n / 128) / max(act.sum(), 1)
    print(f"   stats (cycles per CTA, mean over {int(act.sum())} CTAs, ~{tiles:.2f} tiles each): total {st[act,4].mean():.0f}  wait A(head) {st[act,0].mean():.0f}  wait A(body) {st[act,1].mean():.0f}  wait W {st[act,2].mean():.0f}  producer wait-empty {st[act,3].mean():.0f}", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.pack_weights(flat_d, out=packed)
e1.record(); torch.cuda.synchronize()
print("pack ms", e0.elapsed_time(e1) / 10)

# ---- the reference's own code path (stock PyTorch) on this GPU ----
from oracle.torch_reference import RefR2L, embed, sample
ref = RefR2L().load_flat(flat).to(dev)
zt = torch.from_numpy(orc.sampler_z_vals(2.0, 6.0)).to(dev)
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32; torch.backends.cudnn.allow_tf32 = tf32
    for n in (4096, 160000):
        o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev)
        with torch.no_grad():
            for _ in range(3): r = ref(embed(sample(o, d, zt)))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): r = ref(embed(sample(o, d, zt)))
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ours = ops.forward(packed, rays_o=o, rays_d=d, z_vals=zt.tolist())
        print(f"stock torch fwd N={n} tf32={tf32}: {ms:.3f} ms {n / ms / 1e3:.2f} Mrays/s ; max rel diff ours vs torch-gpu {float(((ours - r).abs() / r.abs()).max()):.3e}", flush=True)
    # fwd+bwd
    n = 4096
    o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); tgt = torch.rand(n, 3, device=dev)
    for _ in range(3):
        ref.zero_grad(); ((ref(embed(sample(o, d, zt))) - tgt) ** 2).mean().backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ref.zero_grad(); ((ref(embed(sample(o, d, zt))) - tgt) ** 2).mean().backward()
    e1.record(); torch.cuda.synchronize()
    print(f"stock torch fwd+bwd N=4096 tf32={tf32}: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
