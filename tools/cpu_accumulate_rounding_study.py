"""CPU study of the tensor core's accumulator rounding (numpy model: oracle/split_emulation.py, mma_linear): max relative error
of the RGB output against the reference's own fp32 RGB on the golden 200-ray batch, for the ways a Linear's 48 (body) / 192
(head) instructions can be accumulated.  The model rounds the fp32 accumulator TOWARD ZERO after every K = 16 instruction - the
behaviour measured on B200 (tools/gpu_accum_calibrate.py: a fresh GEMM comes out short by a few ulp on average, and the mean
signed error of the forward crosses zero for a debias of ~12 x 2^-24).

    python tools/cpu_accumulate_rounding_study.py

Results (seed-0 weights, tests/golden/r2l_seed0.npz; max relative error / mean signed / rms):

    in-place residual accumulate, big term first, truncating  (round 1; single + pair forms)   6.3e-05   -6.1e-06   1.6e-05
    in-place residual accumulate, small terms first, truncating                                6.2e-05   -5.8e-06   1.6e-05
    fresh accumulator + fp32 residual add, big term first, truncating                          7.3e-06   -1.1e-06   2.2e-06
    fresh accumulator + fp32 residual add, small terms first, truncating                       6.3e-06   -8.4e-07   1.8e-06
      + debias eps_body 12 x 2^-24, eps_head 32 x 2^-24  (the half form today)                 3.7e-06   -2.1e-07   1.0e-06
    fresh accumulator, small terms first, ROUND-TO-NEAREST accumulator (not offered)           9.5e-07   -1.2e-08   2.3e-07

The GPU measured 3.7e-5 (round 1, bf16 planes) and 4.3e-5 (round 2, fp16 planes) for the first row's arithmetic and 2.35e-6
for the fifth: the accumulator's rounding, not the operand format, set the forward error - and a residual stream that lives in
the accumulator takes a truncation from every one of its 43 x 48 instructions."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import split_emulation as se
from r2l_b200.nerf_raybased import init_flat_params

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "r2l_seed0.npz")))
flat = init_flat_params(0).numpy()
x, ref = g["x_embed"], g["rgb"]
rows = (("in-place residual accumulate, big term first, truncating   (round 1 / single + pair forms)", dict(residual="in_place", order="big_first")),
        ("in-place residual accumulate, small terms first, truncating", dict(residual="in_place", order="small_first")),
        ("fresh accumulator + fp32 residual add, big term first, truncating", dict(residual="fresh", order="big_first")),
        ("fresh accumulator + fp32 residual add, small terms first, truncating", dict(residual="fresh", order="small_first")),
        ("  + debias eps_body 12 x 2^-24, eps_head 32 x 2^-24           (half form today)", dict(residual="fresh", order="small_first", eps_body=12 / 2 ** 24, eps_head=32 / 2 ** 24)),
        ("fresh accumulator, small terms first, ROUND-TO-NEAREST accumulator (what a GPU does not offer)", dict(residual="fresh", order="small_first", rounding="rn")))
for name, kw in rows:
    rgb = se.r2l_forward_mma(flat, x, **kw)
    rel = (rgb.astype(np.float64) - ref) / np.abs(ref)
    print(f"{name:100s} max {np.abs(rel).max():.2e}   mean signed {rel.mean():+.2e}   rms {np.sqrt((rel ** 2).mean()):.2e}", flush=True)
