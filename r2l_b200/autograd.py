"""torch.autograd bridge: `loss.backward()` through NeRF_v3_2 runs the fused backward (r2l_backward)."""
from __future__ import annotations

import torch

from . import ops


class _R2LFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flat, model, inputs):
        rgb, tctx = ops.forward_train(model.packed_weights(), **inputs)
        ctx.tctx, ctx.model = tctx, model
        return rgb

    @staticmethod
    def backward(ctx, grad_rgb):
        model = ctx.model
        grads = ops.backward(model.packed_weights(), ctx.tctx, grad_rgb.contiguous())
        ctx.tctx = None  # release the saved operand images
        return grads, None, None


def r2l_apply(model, inputs):
    return _R2LFunction.apply(model.flat, model, inputs)
