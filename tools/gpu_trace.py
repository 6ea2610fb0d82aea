"""Per-layer time line of the forward chain kernel (debug stamps).  Usage: gpu_trace.py [infer|train] [form 0|1|2]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import _lib, ops
from r2l_b200.nerf_raybased import init_flat_params
from oracle import r2l_oracle as orc
dev = torch.device("cuda:0")
packed = ops.pack_weights(init_flat_params(0).to(dev))
n = 4096
o = torch.randn(n, 3, device=dev) * 0.5; d = torch.randn(n, 3, device=dev); z = orc.sampler_z_vals(2.0, 6.0).tolist()
mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
form = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # launch form of chain.cu: 0 single, 1 pair, 2 half
_lib.lib().r2l_set_pair_mode(form)
cta = 2 if form else 3                                # a leader CTA (the MMA thread's stamps live there)
tensor = 3078 if form == 2 else 6151
run = (lambda: ops.forward(packed, rays_o=o, rays_d=d, z_vals=z)) if mode == "infer" else (lambda: ops.forward_train(packed, rays_o=o, rays_d=d, z_vals=z))
for _ in range(3): run()
trace = torch.zeros(148 * 5 * 96 + 360, dtype=torch.int64, device=dev)
_lib.lib().r2l_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
run(); torch.cuda.synchronize()
_lib.lib().r2l_debug_set_trace(None)
print("mode:", mode, "form:", form)
t = trace[:148 * 5 * 96].view(148, 5, 96).cpu().numpy().astype(np.int64)[cta]
start, issued, accdone, pub0, epidone = t
L = 87
print("layer: MMA-start  issued-at  acc-complete  first-publish  epilogue-done   (cycles relative to layer-1 start)")
base = start[1]
for l in (1, 2, 3, 10, 11, 40, 41, 85, 86):
    print(l, start[l] - base, issued[l] - base, accdone[l] - base, pub0[l] - base, epidone[l] - base)
body = np.arange(1, 86)
print("mean over body layers:")
print("  layer period (start l+1 - start l)        ", np.mean(start[body + 1] - start[body]))
print("  MMA start -> accumulator seen complete     ", np.mean(accdone[body] - start[body]), f" (tensor work of a layer = {tensor})")
print("  MMA start -> all 48 issued                 ", np.mean(issued[body] - start[body]))
print("  acc complete -> first k-step published     ", np.mean(pub0[body] - accdone[body]))
print("  first publish -> next layer's first MMA    ", np.mean(start[body + 1] - pub0[body]))
print("  acc complete -> epilogue done (4 chunks)   ", np.mean(epidone[body] - accdone[body]))
for name, sel in (("odd layers (first Linear of a block: H -> hidden operand)", body[body % 2 == 1]), ("even layers (second Linear: result joins the stream)", body[body % 2 == 0])):
    print(name + ": period", np.mean(start[sel + 1] - start[sel]), " MMA start->acc", np.mean(accdone[sel] - start[sel]),
          " acc->first publish", np.mean(pub0[sel] - accdone[sel]), " publish->next MMA", np.mean(start[sel + 1] - pub0[sel]),
          " acc->epilogue done", np.mean(epidone[sel] - accdone[sel]))

