"""CPU emulation of the kernels' split arithmetic (TEST INFRASTRUCTURE ONLY, same rules as r2l_oracle.py).

The chain kernels compute every Linear as three fp16 tensor-core products with fp32 accumulation,
    a . w  ~=  (a_hi . W_hi + a_lo . W_hi + a_hi . W_lo) / s,    W = s w,  x_hi = fp16(x), x_lo = fp16(x - x_hi)
with the power-of-two weight scale s = WEIGHT_SCALE (ptx.cuh: kWeightScale; chain.cu; DESIGN.md section 4 "Precision"),
and the backward pass runs on loss_scale * dL/d(.) (dw.cu: r2l_bwd_prep_kernel).  This module restates that arithmetic in
numpy so that the precision claims - three products meet the 1e-3 bar of BASELINE.json with a wide margin, one product
does not; the gradients are as accurate as plain fp32 arithmetic - are checked on the CPU, independently of the GPU
(tests/test_oracle.py).  `fmt="bf16"` keeps round 1's format for comparison (tools/cpu_*_precision_study.py).
Network structure: model/nerf_raybased.py:443-465,:539-544 as in r2l_oracle.r2l_forward."""
from __future__ import annotations

import numpy as np

from . import r2l_oracle as orc

WEIGHT_SCALE = 64.0     # ptx.cuh: kWeightScale


def to_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned as float32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)
    return rounded.view(np.float32)


def to_fp16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to IEEE half (gradual underflow, as cvt.rn.f16x2.f32 does), returned as float32."""
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


_ROUND = {"fp16": to_fp16, "bf16": to_bf16}


def split(x: np.ndarray, fmt: str = "fp16"):
    rnd = _ROUND[fmt]
    hi = rnd(x)
    return hi, rnd(np.asarray(x, np.float32) - hi)


def split_matmul(a: np.ndarray, b: np.ndarray, terms: int = 3, fmt: str = "fp16", b_scale: float = 1.0) -> np.ndarray:
    """a [M,K] @ b [K,N] with `terms` of the three split products (1: hi.hi only), fp32 accumulation; b is packed as
    b_scale * b and the result divided by b_scale (exact for powers of two)."""
    s = np.float32(b_scale)
    a_hi, a_lo = split(a, fmt)
    b_hi, b_lo = split(np.asarray(b, np.float32) * s, fmt)
    out = a_hi @ b_hi
    if terms >= 2:
        out = out + a_lo @ b_hi
    if terms >= 3:
        out = out + a_hi @ b_lo
    return (out / s).astype(np.float32)


def split_linear(a: np.ndarray, w: np.ndarray, terms: int = 3, fmt: str = "fp16", w_scale: float | None = None) -> np.ndarray:
    """a [N,K] @ w[O,K]^T as the chain kernels do it."""
    if w_scale is None:
        w_scale = WEIGHT_SCALE if fmt == "fp16" else 1.0
    return split_matmul(a, np.asarray(w).T, terms, fmt, w_scale)


def r2l_forward_split(flat: np.ndarray, x: np.ndarray, terms: int = 3, fmt: str = "fp16", w_scale: float | None = None) -> np.ndarray:
    """r2l_oracle.r2l_forward with every body/head product done as the kernels do it (the 3-wide tail stays fp32, as in
    the kernels' CUDA-core tail)."""
    p = orc.unflatten_params(flat.astype(np.float32))
    x = x.astype(np.float32)
    lin = lambda a, w: split_linear(a, w, terms, fmt, w_scale)
    h = np.maximum(lin(x, p["head_w"]) + p["head_b"], 0)
    z = h
    for k in range(orc.N_BLOCKS):
        (w1, b1), (w2, b2) = p["body"][2 * k], p["body"][2 * k + 1]
        a = np.maximum(lin(z, w1) + b1, 0)
        z = (lin(a, w2) + b2) + z
    zf = z + h
    return orc.sigmoid(zf @ p["tail_w"].T + p["tail_b"])


def loss_scale_for(grad_rgb: np.ndarray) -> float:
    """dw.cu: r2l_bwd_prep_kernel - the power of two S with S * max |grad_rgb| in [2^9, 2^10)."""
    m = float(np.max(np.abs(grad_rgb)))
    if not (0.0 < m < 3.0e38):
        return 1.0
    _, e = np.frexp(np.float32(m))
    return float(np.ldexp(1.0, int(np.clip(10 - int(e), -100, 100))))


def r2l_grads_split(flat: np.ndarray, x: np.ndarray, target: np.ndarray, fmt: str = "fp16", w_scale: float | None = None,
                    loss_scale: float | None = None):
    """(rgb, flat gradient [NUM_PARAMS] float64) of img2mse through the kernels' arithmetic: forward as above; backward
    chain and weight gradients as chain.cu (kBwd) / dw.cu build them - every product three split terms, dY carrying the
    loss scale, tail gradients in fp32 on the unscaled d logit (r2l_tail_grad_kernel)."""
    if w_scale is None:
        w_scale = WEIGHT_SCALE if fmt == "fp16" else 1.0
    p = orc.unflatten_params(flat.astype(np.float32))
    x = x.astype(np.float32)
    n = x.shape[0]
    lin = lambda a, w: split_linear(a, w, 3, fmt, w_scale)              # a @ w^T, weights packed scaled
    lin_t = lambda a, w: split_matmul(a, w, 3, fmt, w_scale)           # a @ w   (the transposed images of the backward)
    outer = lambda dy, xx: split_matmul(dy.T, xx, 3, fmt, 1.0)         # dW = dY^T X over the ray axis (dw.cu)
    h = np.maximum(lin(x, p["head_w"]) + p["head_b"], 0)
    z, zs, acts = h, [], []
    for k in range(orc.N_BLOCKS):
        (w1, b1), (w2, b2) = p["body"][2 * k], p["body"][2 * k + 1]
        a = np.maximum(lin(z, w1) + b1, 0)
        zs.append(z)
        acts.append(a)
        z = (lin(a, w2) + b2) + z
    zf = z + h
    rgb = orc.sigmoid(zf @ p["tail_w"].T + p["tail_b"])
    grad_rgb = ((2.0 / (3 * n)) * (rgb - target.astype(np.float32))).astype(np.float32)
    s = np.float32(loss_scale_for(grad_rgb) if loss_scale is None else loss_scale)
    g = np.zeros(orc.NUM_PARAMS, np.float64)
    dl = (grad_rgb * rgb * (1 - rgb)).astype(np.float32)
    g[orc.OFF_TAIL_W:orc.OFF_TAIL_B] = (dl.T @ zf).reshape(-1)
    g[orc.OFF_TAIL_B:] = dl.sum(0)
    gz = ((dl * s) @ p["tail_w"]).astype(np.float32)
    g43 = gz.copy()
    for k in range(orc.N_BLOCKS - 1, -1, -1):
        w1, w2 = p["body"][2 * k][0], p["body"][2 * k + 1][0]
        o1 = orc.OFF_BODY + (2 * k) * orc.LINEAR_STRIDE
        o2 = o1 + orc.LINEAR_STRIDE
        g[o2:o2 + 65536] = outer(gz, acts[k]).reshape(-1) / s
        g[o2 + 65536:o2 + 65792] = gz.sum(0) / s
        dh = lin_t(gz, w2) * (acts[k] > 0)
        g[o1:o1 + 65536] = outer(dh, zs[k]).reshape(-1) / s
        g[o1 + 65536:o1 + 65792] = dh.sum(0) / s
        gz = gz + lin_t(dh, w1)
    dhead = (gz + g43) * (h > 0)
    g[orc.OFF_HEAD_W:orc.OFF_HEAD_B] = outer(dhead, x).reshape(-1) / s
    g[orc.OFF_HEAD_B:orc.OFF_BODY] = dhead.sum(0) / s
    return rgb, g


# ------------------------------------------------------------------------------------------------
# accumulation model: the tensor core rounds its fp32 accumulator TOWARD ZERO after every tcgen05.mma (measured on B200:
# tools/gpu_accum_calibrate.py, DESIGN.md section 4 "Precision").  The functions below issue a Linear as the kernels do -
# K = 16 per instruction, 64-wide K chunks, three split products per chunk - and round the accumulator after every instruction,
# so that the consequences (in-place residual accumulation vs a fresh accumulator per GEMM, issue order of the split terms) can
# be studied on the CPU (tools/cpu_accumulate_rounding_study.py, tests/test_oracle.py).
# ------------------------------------------------------------------------------------------------
def round_f32(x64: np.ndarray, mode: str = "rz") -> np.ndarray:
    """float64 -> float32, mode "rz" = toward zero (what the accumulator does), "rn" = to nearest even."""
    y = np.asarray(x64, np.float64).astype(np.float32)
    if mode == "rn":
        return y
    over = np.abs(y.astype(np.float64)) > np.abs(x64)               # rounded away from zero: step back by one ulp
    return np.where(over, np.nextafter(y, np.float32(0)), y).astype(np.float32)


def mma_linear(a: np.ndarray, w: np.ndarray, acc0: np.ndarray | None = None, order: str = "small_first", rounding: str = "rz",
               w_scale: float = WEIGHT_SCALE) -> np.ndarray:
    """The raw fp32 accumulator after a [N,K] x w[O,K]^T Linear issued as K/16 x 3 instructions on fp16 hi/lo planes (weights
    packed as w_scale * w), starting from acc0 (None = a fresh accumulator).  order: per 64-wide chunk, "small_first" =
    a_lo W_hi, a_hi W_lo, a_hi W_hi (round 2's half form) or "big_first" = a_hi W_hi, a_lo W_hi, a_hi W_lo (round 1).  The
    caller divides by w_scale (exact) and adds bias / residual."""
    a_hi, a_lo = split(a, "fp16")
    w_hi, w_lo = split(np.asarray(w, np.float32) * np.float32(w_scale), "fp16")
    pairs = {"small_first": ((a_lo, w_hi), (a_hi, w_lo), (a_hi, w_hi)), "big_first": ((a_hi, w_hi), (a_lo, w_hi), (a_hi, w_lo))}[order]
    acc = np.zeros((a.shape[0], w.shape[0]), np.float32) if acc0 is None else np.asarray(acc0, np.float32).copy()
    k_total = a.shape[1]
    for c0 in range(0, k_total, 64):
        for pa, pw in pairs:
            for k0 in range(c0, min(c0 + 64, k_total), 16):
                part = pa[:, k0:k0 + 16].astype(np.float64) @ pw[:, k0:k0 + 16].astype(np.float64).T     # exact products, wide sum
                acc = round_f32(acc.astype(np.float64) + part, rounding)
    return acc


def r2l_forward_mma(flat: np.ndarray, x: np.ndarray, residual: str = "fresh", order: str = "small_first", rounding: str = "rz",
                    eps_body: float = 0.0, eps_head: float = 0.0) -> np.ndarray:
    """The forward pass with the accumulation model above.  residual = "in_place": the second Linear of every block accumulates
    straight onto the residual stream kept in the accumulator (round 1, and the single / pair forms today); "fresh": every
    GEMM starts from zero and the epilogue adds its result to the stream in fp32, round to nearest (the half form).  eps_*:
    the (1 + eps) debias the half form applies to every accumulator it reads."""
    p = orc.unflatten_params(flat.astype(np.float32))
    s = np.float32(WEIGHT_SCALE)
    x = x.astype(np.float32)
    kw = dict(order=order, rounding=rounding)
    fb, fh = np.float32(1.0 + eps_body) / s, np.float32(1.0 + eps_head) / s
    h = np.maximum(mma_linear(x, p["head_w"], **kw) * fh + p["head_b"], 0).astype(np.float32)
    z = h
    for k in range(orc.N_BLOCKS):
        (w1, b1), (w2, b2) = p["body"][2 * k], p["body"][2 * k + 1]
        a = np.maximum(mma_linear(z, w1, **kw) * fb + b1, 0).astype(np.float32)
        if residual == "in_place":
            z = (mma_linear(a, w2, acc0=z * s, **kw) / s + b2).astype(np.float32)      # the stream lives in the accumulator
        else:
            z = ((mma_linear(a, w2, **kw) * fb + b2) + z).astype(np.float32)
    return orc.sigmoid((z + h) @ p["tail_w"].T + p["tail_b"])
