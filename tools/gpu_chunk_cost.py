"""Cost of completing the gradient buffer in chunks (r2l_backward_chunked) on ONE GPU, without any collective: forward_train +
backward at 4096 rays for several (split_layers, reserve_sms).  Usage: python tools/gpu_chunk_cost.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import r2l_oracle as orc
from r2l_b200 import ops
from r2l_b200.nerf_raybased import init_flat_params
dev = torch.device("cuda:0")
packed = ops.pack_weights(init_flat_params(0).to(dev))
n = 4096
o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); t = torch.rand(n, 3, device=dev)
z = orc.sampler_z_vals(2.0, 6.0).tolist(); gr = torch.empty(ops.NUM_PARAMS, device=dev)
nf, nb_, nw = ops.train_buffer_bytes(n)
fs, bs, ws = (torch.empty(k, dtype=torch.uint8, device=dev) for k in (nf, nb_, nw))
for split, sms in (((), 0), ((43,), 0), ((43,), 16), ((64, 43, 21), 0), ((64, 43, 21), 16), ((70, 56, 43, 30, 16), 8)):
    def step():
        rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z, fwd_saved=fs, workspace=ws)
        ops.backward(packed, ctx, (2.0 / (3 * n)) * (rgb - t), gr, bwd_saved=bs, workspace=ws, split_layers=list(split), reserve_sms=sms)
    for _ in range(3): step()
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        step()
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"split {split} reserve_sms {sms}: fwd+bwd {e0.elapsed_time(e1) / 50:.4f} ms", flush=True)
