"""CPU study of the GRADIENT precision of split operand formats (numpy emulation of the kernels' forward + backward with
three split products per GEMM; flat-buffer Frobenius error against the fp64 gradients of oracle/r2l_oracle.py).
Round-1 result at 1024 lego-pose rays, seed-0 weights:

    fp32 products (the reference's arithmetic)   2.5e-06
    bf16 x3 (the kernels today)                  1.9e-03   (the GPU measures 1.0e-3 at 1000 rays, 5-6e-4 at 4096)
    fp16 x3 with a 2^18 loss scale               3.4e-04
    fp16 x3 without loss scale                   3.5e-03   (dY underflows fp16)

The gradient error is ~170x the forward error for either format (ReLU masks and dY both inherit it), so the 6x more
accurate fp16 planes carry over to the gradients.  Usage: python tools/cpu_gradient_precision_study.py [n_rays]"""
import importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import r2l_oracle as orc, split_emulation as se
from r2l_b200.nerf_raybased import init_flat_params
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
flat = init_flat_params(0).numpy()
bf16 = se.to_bf16
fp16 = lambda x: np.asarray(x, np.float32).astype(np.float16).astype(np.float32)

def make_prod(rnd):
    def split(x):
        hi = rnd(x); return hi, rnd(x.astype(np.float32) - hi)
    def prod(a, b):     # a [M,K] @ b [K,N], three split products, fp32 accumulate
        a_hi, a_lo = split(a); b_hi, b_lo = split(b)
        return (a_hi @ b_hi + a_lo @ b_hi + a_hi @ b_lo).astype(np.float32)
    return prod

def grads(prod, x, target, scale=1.0):
    p = orc.unflatten_params(flat); x = x.astype(np.float32); N = x.shape[0]
    h = np.maximum(prod(x, p["head_w"].T) + p["head_b"], 0); z = h; zs, as_ = [], []
    for k in range(orc.N_BLOCKS):
        (w1, b1), (w2, b2) = p["body"][2 * k], p["body"][2 * k + 1]
        a = np.maximum(prod(z, w1.T) + b1, 0); zs.append(z); as_.append(a)
        z = (prod(a, w2.T) + b2) + z
    zf = z + h
    rgb = orc.sigmoid(zf @ p["tail_w"].T + p["tail_b"])
    g = np.zeros(orc.NUM_PARAMS, np.float64)
    dl = ((2.0 / (3 * N)) * (rgb - target) * rgb * (1 - rgb)).astype(np.float32) * np.float32(scale)
    g[orc.OFF_TAIL_W:orc.OFF_TAIL_B] = (dl.T.astype(np.float64) @ zf.astype(np.float64)).reshape(-1) / scale
    g[orc.OFF_TAIL_B:] = dl.sum(0) / scale
    gz = (dl @ p["tail_w"]).astype(np.float32); g43 = gz.copy()
    for k in range(orc.N_BLOCKS - 1, -1, -1):
        (w1, b1), (w2, b2) = p["body"][2 * k], p["body"][2 * k + 1]
        o2 = orc.OFF_BODY + (2 * k + 1) * orc.LINEAR_STRIDE; o1 = orc.OFF_BODY + (2 * k) * orc.LINEAR_STRIDE
        g[o2:o2 + 65536] = prod(gz.T, as_[k]).reshape(-1) / scale; g[o2 + 65536:o2 + 65792] = gz.sum(0) / scale
        da = prod(gz, w2); dh = da * (as_[k] > 0)
        g[o1:o1 + 65536] = prod(dh.T, zs[k]).reshape(-1) / scale; g[o1 + 65536:o1 + 65792] = dh.sum(0) / scale
        gz = gz + prod(dh, w1)
    dhead = (gz + g43) * (h > 0)
    g[orc.OFF_HEAD_W:orc.OFF_HEAD_B] = prod(dhead.T, x).reshape(-1) / scale; g[orc.OFF_HEAD_B:orc.OFF_BODY] = dhead.sum(0) / scale
    return g

z = orc.sampler_z_vals(2.0, 6.0)
ro, rd, tg = bench.synthetic_rays(n, 0)
x = orc.positional_embed(orc.sample_train(ro, rd, z, None))
t0 = time.time()
_, g64, _, _ = orc.r2l_loss_and_grads(flat.astype(np.float64), x.astype(np.float64), tg.astype(np.float64))
print("fp64 grads", time.time() - t0, "s", flush=True)
fp32prod = lambda a, b: (a.astype(np.float32) @ b.astype(np.float32))
for name, prod, scale in (("fp32 (reference arithmetic)", fp32prod, 1.0), ("bf16 x3", make_prod(bf16), 1.0), ("fp16 x3, loss scale 2^18", make_prod(fp16), 2.0 ** 18),
                          ("fp16 x3, no loss scale", make_prod(fp16), 1.0)):
    t0 = time.time(); g = grads(prod, x, tg, scale)
    print(f"{name:28s} flat Frobenius rel err vs fp64: {np.linalg.norm(g - g64) / np.linalg.norm(g64):.2e}   ({time.time() - t0:.0f}s)", flush=True)
