"""CPU study of the GRADIENT precision of split operand formats (oracle/split_emulation.py: numpy emulation of the kernels'
forward + backward with three split products per GEMM; errors against the fp64 gradients of oracle/r2l_oracle.py).
Results at 1024 lego-pose rays, seed-0 weights (flat-buffer Frobenius / worst tensor):

    fp32 products (plain numpy matmuls)                  2.5e-06 / 1.3e-05
    bf16 x3 (round 1's kernels)                          1.9e-03 / 4.7e-03   (the GPU measured 1.0e-3 at 1000 rays)
    fp16 x3, loss scale, weights unscaled                3.4e-04 / 8.5e-04
    fp16 x3, loss scale, weights x 16 or x 64 (today)    2.5e-06 / 1.3e-05   = the fp32 row, tensor for tensor
    fp16 x3 without a loss scale                         3.5e-03             (dY underflows fp16)

Why the weight scale matters: default-initialised weights are ~2^-5, so the lo plane of an UNSCALED weight (<= 2^-11 |w|)
is an fp16 subnormal and carries only 2^-25 absolute, i.e. ~2^-19 relative precision; times 64 it is a normal number and
hi + lo carries 22 bits.  The result does not depend on the loss scale within a factor of 2^8 either way.
Usage: python tools/cpu_gradient_precision_study.py [n_rays]"""
import importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import r2l_oracle as orc, split_emulation as se
from r2l_b200.nerf_raybased import init_flat_params
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
flat = init_flat_params(0).numpy()
z = orc.sampler_z_vals(2.0, 6.0)
ro, rd, tg = bench.synthetic_rays(n, 0)
x = orc.positional_embed(orc.sample_train(ro, rd, z, None))
_, g64, _, _ = orc.r2l_loss_and_grads(flat.astype(np.float64), x.astype(np.float64), tg.astype(np.float64))
_, g32, _, _ = orc.r2l_loss_and_grads(flat.astype(np.float32), x.astype(np.float32), tg.astype(np.float32))
bounds = [(orc.OFF_HEAD_W, orc.OFF_HEAD_B), (orc.OFF_HEAD_B, orc.OFF_BODY)]
for l in range(2 * orc.N_BLOCKS):
    o = orc.OFF_BODY + l * orc.LINEAR_STRIDE
    bounds += [(o, o + 65536), (o + 65536, o + 65792)]
bounds += [(orc.OFF_TAIL_W, orc.OFF_TAIL_B), (orc.OFF_TAIL_B, orc.NUM_PARAMS)]
per_tensor = lambda g: np.array([np.linalg.norm(g[a:b] - g64[a:b]) / np.linalg.norm(g64[a:b]) for a, b in bounds])
ref = per_tensor(g32.astype(np.float64))
print(f"{'fp32 numpy oracle':44s} flat {np.linalg.norm(g32 - g64) / np.linalg.norm(g64):.2e}  worst tensor {ref.max():.2e}")
auto = se.loss_scale_for(np.float32(2.0 / (3 * n)) * np.ones(1, np.float32))
for name, kw in (("bf16 x3", dict(fmt="bf16", loss_scale=1.0)), ("fp16 x3, w x 1, auto loss scale", dict(w_scale=1.0)),
                 ("fp16 x3, w x 64, auto loss scale (kernels)", dict()), ("fp16 x3, w x 16", dict(w_scale=16.0)),
                 ("fp16 x3, w x 64, loss scale / 256", dict(loss_scale=auto / 256)), ("fp16 x3, w x 64, loss scale x 256", dict(loss_scale=auto * 256)),
                 ("fp16 x3, w x 64, no loss scale", dict(loss_scale=1.0))):
    t0 = time.time()
    _, g = se.r2l_grads_split(flat, x, tg, **kw)
    pt = per_tensor(g)
    print(f"{name:44s} flat {np.linalg.norm(g - g64) / np.linalg.norm(g64):.2e}  worst tensor {pt.max():.2e}  worst ratio to the fp32 row "
          f"{np.max(pt / ref):.2f} (median {np.median(pt / ref):.2f})  ({time.time() - t0:.0f}s)", flush=True)
