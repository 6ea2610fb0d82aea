"""Adam on flat fp32 CUDA parameters in one kernel.

Drop-in for `torch.optim.Adam(model.parameters(), lr=..., betas=(0.9, 0.999))` (main.py:465) when every parameter is a
contiguous fp32 CUDA tensor — which is the case for NeRF_v3_2 here (one flat parameter).  Same update rule and state
names (`step`, `exp_avg`, `exp_avg_sq`), so optimizer checkpoints stay readable; the learning-rate schedule of
main.py:1181-1195 keeps working because the kernel reads `param_group['lr']` at every step."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FlatAdam: parameters must be contiguous float32 CUDA tensors (no CPU fallback)")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                with torch.cuda.device(p.device):
                    _lib.check(L.r2l_adam_step(ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(g.data_ptr()),
                                               ctypes.c_void_p(st["exp_avg"].data_ptr()), ctypes.c_void_p(st["exp_avg_sq"].data_ptr()),
                                               p.numel(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                               int(st["step"]), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                               "r2l_adam_step")
                # the kernel wrote through the raw pointer: tell autograd / the packed-weight cache that p changed
                torch.autograd.graph.increment_version(p)
        return loss
