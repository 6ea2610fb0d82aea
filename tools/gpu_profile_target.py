"""Workload for `ncu --set full`: a few training iterations at 4096 rays with plain launches (no CUDA graph), so that the
chain kernels (forward-train, backward) and the weight-gradient kernel appear as separate profiled launches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import ops
from r2l_b200.nerf_raybased import init_flat_params
from oracle import r2l_oracle as orc
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3       # (compute-sanitizer runs: 1)
packed = ops.pack_weights(init_flat_params(0).to(dev)); z = orc.sampler_z_vals(2.0, 6.0).tolist()
torch.manual_seed(1); o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); t = torch.rand(n, 3, device=dev)
gr = torch.empty(ops.NUM_PARAMS, device=dev)
for _ in range(iters):
    rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z)
    _, g, _ = ops.mse_loss_grad(rgb, t, 2.0 / (3 * n), 1.0 / (3 * n))
    ops.backward(packed, ctx, g, gr)
    out = ops.forward(packed, rays_o=o, rays_d=d, z_vals=z)
torch.cuda.synchronize()
print("done")
