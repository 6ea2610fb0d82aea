"""Teacher NeRF (8x256 MLP with skip and view branch) parameter handling; reference model/nerf_raybased.py:337-401."""
from __future__ import annotations

import torch
import torch.nn as nn


def teacher_layer_shapes(D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,)):
    """(out, in) per nn.Linear in the reference's construction order (:357-375)."""
    shapes = [(W, input_ch)] + [(W, W + input_ch) if i in skips else (W, W) for i in range(D - 1)]
    shapes += [(W // 2, input_ch_views + W)]            # views_linears.0
    shapes += [(W, W), (1, W), (3, W // 2)]              # feature_linear, alpha_linear, rgb_linear
    return shapes


def init_teacher_params(seed=None, **kw):
    """Default-initialised parameters drawn in the reference's order; returns [w0, b0, w1, b1, ...] in
    state_dict order (pts_linears.*, views_linears.0, feature_linear, alpha_linear, rgb_linear)."""
    if seed is not None:
        torch.manual_seed(seed)
    out = []
    for o, i in teacher_layer_shapes(**kw):
        lin = nn.Linear(i, o)
        out += [lin.weight.detach().clone(), lin.bias.detach().clone()]
    return out
