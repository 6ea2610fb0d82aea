"""MMA-thread wait statistics of the forward chain kernel in its three launch forms.  Usage: gpu_stats.py [n_rays]"""
import ctypes, sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import _lib, ops
from r2l_b200.nerf_raybased import init_flat_params
from oracle import r2l_oracle as orc
dev = torch.device("cuda:0"); packed = ops.pack_weights(init_flat_params(0).to(dev)); L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); z = orc.sampler_z_vals(2.0, 6.0).tolist()
for pair in (0, 1, 2):
    L.r2l_set_pair_mode(pair)
    for train in (0, 1):
        run = (lambda: ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z)) if train else (lambda: ops.forward(packed, rays_o=o, rays_d=d, z_vals=z))
        for _ in range(3): run()
        st = torch.zeros(148 * 8, dtype=torch.int64, device=dev); L.r2l_debug_set_stats(ctypes.c_void_p(st.data_ptr()))
        run(); torch.cuda.synchronize(); L.r2l_debug_set_stats(None)
        s = st.view(148, 8).cpu().numpy().astype(float); a = s[:, 4] > 0
        print(f"form={pair} train={train}: MMA-thread total {s[a,4].mean():.0f}  wait A head {s[a,0].mean():.0f}  wait A body {s[a,1].mean():.0f}  wait W {s[a,2].mean():.0f}  producer wait-empty {s[s[:,3]>0,3].mean():.0f}", flush=True)
