"""Sweep of the weight-gradient schedule (r2l_debug_set_dw_schedule): backward time at 4096 rays (dW overlaps the chain)
and at a large batch (dW after the chain), gradients compared with the first schedule.  Usage: gpu_dw_sched.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import _lib, ops
from r2l_b200.nerf_raybased import init_flat_params
from oracle import r2l_oracle as orc
dev = torch.device("cuda:0"); L = _lib.lib()
packed = ops.pack_weights(init_flat_params(0).to(dev)); z = orc.sampler_z_vals(2.0, 6.0).tolist()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(n, sched, reps=10):
    torch.manual_seed(1); o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); t = torch.rand(n, 3, device=dev)
    gr = torch.empty(ops.NUM_PARAMS, device=dev)
    L.r2l_debug_set_dw_schedule(*sched)   # (t1, t2, t3, serial, t4)
    tot_b = tot = 0.0
    for i in range(reps + 3):
        flush.fill_(1)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); rgb, ctx = ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z)
        g = (2.0 / (3 * n)) * (rgb - t); e1.record()
        ops.backward(packed, ctx, g, gr); e2.record(); torch.cuda.synchronize()
        if i >= 3: tot_b += e1.elapsed_time(e2); tot += e0.elapsed_time(e2)
    return tot / reps, tot_b / reps, gr.clone()

det = "det" in sys.argv[1:]
L.r2l_set_deterministic(1 if det else 0)
print("deterministic mode:", det, flush=True)
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
base = None
scheds = [(-1, -1, -1, 0, -1)] if quick else [(-1, -1, -1, 0, -1), (0, 0, 84, 0, 90), (0, 0, 0, 0, 90), (0, 0, 80, 0, 84), (0, 0, 0, 0, 80), (0, 0, 60, 0, 86), (0, 0, 0, 0, 0), (90, 90, 90, 0, 90)]
for sched in scheds:
    ms, ms_b, gr = timed(4096, sched, reps=3 if quick else 10)
    if base is None: base = gr
    print(f"N=4096 sched {sched}: fwd+bwd {ms:.4f} ms, backward {ms_b:.4f} ms, grads vs first rel diff {float((gr - base).norm() / base.norm()):.2e}", flush=True)
for n in ((1000, 18944) if quick else (1000, 2048, 5376, 18944, 98304)):
    ms, ms_b, gr = timed(n, (-1, -1, -1, 0, -1), reps=4)
    print(f"N={n} default schedule: fwd+bwd {ms:.4f} ms, backward {ms_b:.4f} ms ({n / ms / 1e3:.2f} M rays/s), finite={bool(torch.isfinite(gr).all())}", flush=True)
L.r2l_debug_set_dw_schedule(-1, -1, -1, 0, -1)
