"""Does dw.cu overlap with the backward chain?  globaltimer stamps from both kernels."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import _lib, ops
from r2l_b200.nerf_raybased import init_flat_params
from oracle import r2l_oracle as orc
dev = torch.device("cuda:0")
packed = ops.pack_weights(init_flat_params(0).to(dev))
n = 4096
o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); t = torch.rand(n, 3, device=dev)
z = orc.sampler_z_vals(2.0, 6.0).tolist()
for _ in range(3):
    rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z); ops.backward(packed, ctx, (rgb - t) * 1e-4)
rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z); torch.cuda.synchronize()
stats = torch.zeros(148 * 8, dtype=torch.int64, device=dev); times = torch.zeros(148 * 5 * 96 + 90 * 4, dtype=torch.int64, device=dev)
L = _lib.lib()
L.r2l_debug_set_stats(ctypes.c_void_p(stats.data_ptr())); L.r2l_debug_set_trace(ctypes.c_void_p(times.data_ptr()))
ops.backward(packed, ctx, (rgb - t) * 1e-4); torch.cuda.synchronize()
L.r2l_debug_set_stats(None); L.r2l_debug_set_trace(None)
st = stats.view(148, 8).cpu().numpy(); tm = times.cpu().numpy()[148 * 5 * 96:].reshape(90, 4)
chain_end = st[:32, 5].max()
t0 = tm[:, 0].min()
print("dw kernel first CTA start (us rel.):", 0.0, " chain MMA threads end at", (chain_end - t0) / 1e3)
for u in (85, 84, 60, 30, 1, 0, 86, 89):
    print(f"unit {u}: start {(tm[u,0]-t0)/1e3:8.1f}  flag seen {(tm[u,1]-t0)/1e3:8.1f}  end {(tm[u,3]-t0)/1e3:8.1f} us")
print("last dW CTA end:", (tm[:, 3].max() - t0) / 1e3, "us; sum of unit busy times / 90:", np.mean(tm[:, 3] - tm[:, 1]) / 1e3)
