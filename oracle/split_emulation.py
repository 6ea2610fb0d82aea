"""CPU emulation of the kernels' split arithmetic (TEST INFRASTRUCTURE ONLY, same rules as r2l_oracle.py).

The chain kernels compute every Linear as three bf16 tensor-core products with fp32 accumulation,
    a . w  ~=  a_hi . w_hi + a_lo . w_hi + a_hi . w_lo,      x_hi = bf16(x), x_lo = bf16(x - x_hi)
(chain.cu; DESIGN.md section 4 "Precision").  This module restates that arithmetic in numpy so that the precision claim
- three products meet the 1e-3 bar of BASELINE.json with a wide margin, one bf16 product does not - is checked on the CPU,
independently of the GPU (tests/test_oracle.py).  Network structure: model/nerf_raybased.py:443-465,:539-544 as in
r2l_oracle.r2l_forward."""
from __future__ import annotations

import numpy as np

from . import r2l_oracle as orc


def to_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned as float32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)
    return rounded.view(np.float32)


def split(x: np.ndarray):
    hi = to_bf16(x)
    return hi, to_bf16(x.astype(np.float32) - hi)


def split_linear(a: np.ndarray, w: np.ndarray, terms: int = 3) -> np.ndarray:
    """a [N,K] @ w[O,K]^T with `terms` of the three split products (1: hi.hi only), fp32 accumulation."""
    a_hi, a_lo = split(a)
    w_hi, w_lo = split(w)
    out = a_hi @ w_hi.T
    if terms >= 2:
        out = out + a_lo @ w_hi.T
    if terms >= 3:
        out = out + a_hi @ w_lo.T
    return out.astype(np.float32)


def r2l_forward_split(flat: np.ndarray, x: np.ndarray, terms: int = 3) -> np.ndarray:
    """r2l_oracle.r2l_forward with every body/head product done as the kernels do it (the 3-wide tail stays fp32, as in
    the kernels' CUDA-core tail)."""
    p = orc.unflatten_params(flat.astype(np.float32))
    x = x.astype(np.float32)
    h = np.maximum(split_linear(x, p["head_w"], terms) + p["head_b"], 0)
    z = h
    for k in range(orc.N_BLOCKS):
        (w1, b1), (w2, b2) = p["body"][2 * k], p["body"][2 * k + 1]
        a = np.maximum(split_linear(z, w1, terms) + b1, 0)
        z = (split_linear(a, w2, terms) + b2) + z
    zf = z + h
    return orc.sigmoid(zf @ p["tail_w"].T + p["tail_b"])
