"""BASELINE config 4: NeRF-teacher pseudo-data generation at 400x400 (create_data.py path): coarse 64 + fine 192 samples,
chunk 32768, use_viewdirs, white_bkgd, perturb 0 — seconds per frame / frames per second on one GPU, and the [H*W, 9]
(o | d | rgb) pseudo-data rows the generator writes (utils/create_data.py:820-872)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import nerf_raybased as nb
from r2l_b200 import render as rr
dev = torch.device("cuda:0"); nb.device = dev
torch.manual_seed(0)
coarse = nb.NeRF(8, 256, 63, 27, 4, [4], True).to(dev); fine = nb.NeRF(8, 256, 63, 27, 4, [4], True).to(dev)
embed_fn, _ = nb.get_embedder(10, 0); embeddirs_fn, _ = nb.get_embedder(4, 0)
query = lambda inputs, viewdirs, network_fn: nb.run_network(inputs, viewdirs, network_fn, embed_fn, embeddirs_fn, netchunk=1024 * 64)
H = W = 400; focal = 555.5555155968841
ps = nb.PointSampler(H, W, focal, 16, 2.0, 6.0)
c2w = torch.tensor([[-0.9, 0.2, -0.3, -1.3], [-0.4, -0.5, 0.7, 3.0], [0.0, 0.8, 0.5, 2.2]], device=dev)
kw = dict(network_fn=coarse, network_query_fn=query, N_samples=64, N_importance=128, network_fine=fine, white_bkgd=True, perturb=0., raw_noise_std=0.)

def frame():
    rays_o, rays_d = ps._pose_rays(c2w)
    rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
    rgb, disp, acc, _ = rr.render(H, W, focal, chunk=1024 * 32, rays=(rays_o, rays_d), ndc=False, near=2., far=6., use_viewdirs=True, **kw)
    return torch.cat([rays_o, rays_d, rgb], -1)     # the rows create_data.py appends to `data`

with torch.no_grad():
    for _ in range(2): d9 = frame()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): d9 = frame()
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"teacher frame 400x400 (64 + 192 samples/ray): {ms:.1f} ms = {1e3 / ms:.2f} frames/s = {H * W / ms / 1e3:.2f} M rays/s; "
      f"{H * W * 256 * 2 * 593408 / ms / 1e9:.0f} TFLOP/s algorithmic; rows {tuple(d9.shape)} finite={bool(torch.isfinite(d9).all())}")
