"""GPU diagnostic: teacher NeRF fused kernel parity + throughput, raw2outputs throughput."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2l_oracle as orc
from r2l_b200 import ops
from r2l_b200 import nerf_raybased as nb
dev = torch.device("cuda:0"); nb.device = dev
t = dict(np.load(os.path.join(ROOT, "tests/golden/teacher_seed0.npz")))
torch.manual_seed(0)
net = nb.NeRF(8, 256, 63, 27, 4, [4], True).to(dev)
pts, vd = torch.from_numpy(t["pts"]).to(dev), torch.from_numpy(t["viewdirs"]).to(dev)
raw = net.query(pts, vd); torch.cuda.synchronize()
print("teacher golden: max |err| / max |raw| =", float(np.abs(raw.cpu().numpy() - t["raw"]).max() / np.abs(t["raw"]).max()), flush=True)
params = [p.detach().cpu().numpy() for p in net.parameters()]
for n, s in ((3, 64), (130, 192)):
    torch.manual_seed(n); p_ = torch.randn(n, s, 3) * 1.5; v_ = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
    r = net.query(p_.to(dev), v_.to(dev)).cpu().numpy(); ref = orc.run_network(p_.numpy(), v_.numpy(), params)
    print(f"teacher n={n} s={s}: max err / max |ref| = {np.abs(r - ref).max() / np.abs(ref).max():.3e}", flush=True)
for n, s in ((32768, 64), (32768, 192)):
    p_ = torch.randn(n, s, 3, device=dev); v_ = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
    for _ in range(2): net.query(p_, v_)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = net.query(p_, v_); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"teacher query {n}x{s}: {ms:.2f} ms  {n * s / ms / 1e3:.1f} Mpts/s  {n * s * 2 * 593408 / ms / 1e9:.0f} TFLOP/s algorithmic", flush=True)
    z = torch.sort(torch.rand(n, s, device=dev) * 4 + 2, -1).values; d = torch.randn(n, 3, device=dev)
    for _ in range(2): ops.raw2outputs(r, z, d, True)
    torch.cuda.synchronize()
    e0.record(); o = ops.raw2outputs(r, z, d, True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); byts = n * (4 * (5 * s + 3) + 4 * (6 + s))
    print(f"raw2outputs {n}x{s}: {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s algorithmic", flush=True)
