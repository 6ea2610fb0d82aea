"""Data-parallel plumbing: one process per GPU, rays sharded, ONE all-reduce of the flat gradient per step.

Replaces the reference's single-process nn.DataParallel (main.py:37-42, :472-479), which re-broadcasts all 23.7 MB of
parameters every forward and reduces gradients to GPU 0.  Rays are independent, so the forward needs no collective; the
backward needs exactly one sum over ranks of the 5,917,187-float gradient buffer.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous ray range [lo, hi) of `rank`; sizes differ by at most one ray."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def mse_grad_rgb(rgb: torch.Tensor, target: torch.Tensor, n_global: int, lw_rgb: float = 1.0) -> torch.Tensor:
    """d/drgb of img2mse over the GLOBAL batch (main.py:1377): summing the per-rank parameter gradients built from
    this gives exactly the reference's full-batch gradient, whatever the shard sizes."""
    return (rgb - target) * (2.0 * lw_rgb / (3 * n_global))


def allreduce_flat_grads(grads: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the flat gradient buffer over ranks in place (NCCL on GPUs, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=group)
    return grads


def gather_rgb(rgb_local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """Inference: every rank renders its ray range; rank order concatenation restores the frame."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rgb_local
    world = dist.get_world_size(group)
    sizes = [shard_range(n, r, world) for r in range(world)]
    parts = [torch.empty((hi - lo, 3), dtype=rgb_local.dtype, device=rgb_local.device) for lo, hi in sizes]
    dist.all_gather(parts, rgb_local.contiguous(), group=group) if len({hi - lo for lo, hi in sizes}) == 1 else \
        [dist.broadcast(parts[r] if r != dist.get_rank(group) else parts[r].copy_(rgb_local), src=r, group=group) for r in range(world)]
    return torch.cat(parts, 0)


class _DeviceBlob:
    """A library-owned device buffer exposed to torch through __cuda_array_interface__ (no copy, no ownership)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerDataParallel:
    """The data-parallel step over NVLink / NVSwitch peer memory (csrc/dp.cu; one process per GPU on one node).

    Construction is collective: every rank allocates its block through the C library, the CUDA IPC handles travel through
    the (any-backend) torch.distributed group once, every rank maps the others' blocks.  Afterwards `grads` and `params` are
    float32 CUDA tensors over this rank's buffers and `adam_step` launches the fused reduce-scatter + Adam + all-gather
    kernel; no NCCL call is made per iteration.  Replaces nn.DataParallel's replicate / gather / reduce (main.py:37-42,
    :472-479)."""

    def __init__(self, n_params: int, device, group=None):
        from . import _lib
        L = _lib.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            handle = ctypes.create_string_buffer(int(L.r2l_dp_handle_bytes()))
            _lib.check(L.r2l_dp_create(self.rank, self.world, int(n_params), handle), "r2l_dp_create")
            blobs = [None] * self.world
            dist.all_gather_object(blobs, handle.raw, group=group)
            joined = ctypes.create_string_buffer(b"".join(blobs), len(handle.raw) * self.world)
            _lib.check(L.r2l_dp_connect(joined), "r2l_dp_connect")
            self.grads = torch.as_tensor(_DeviceBlob(L.r2l_dp_grads(), n_params), device=self.device)
            self.params = torch.as_tensor(_DeviceBlob(L.r2l_dp_params(), n_params), device=self.device)
        self.n = int(n_params)
        dist.barrier(group=group)     # every rank has mapped every block before anyone launches the kernel

    def slice_of(self, lo: int, hi: int):
        """The part of the float range [lo, hi) this rank updates."""
        from . import _lib
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.lib().r2l_dp_slice(int(lo), int(hi), ctypes.byref(a), ctypes.byref(b)), "r2l_dp_slice")
        return a.value, b.value

    def adam_step(self, exp_avg, exp_avg_sq, beta1, beta2, eps, hyper_dev, lo: int = 0, hi: int | None = None, slot: int = 0,
                  grid: int = 0, stream=None):
        """The fused kernel on the float range [lo, hi) of the buffers (default: everything), on `stream` (default: current)."""
        from . import _lib
        stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().r2l_dp_adam_step_range(ctypes.c_void_p(exp_avg.data_ptr()), ctypes.c_void_p(exp_avg_sq.data_ptr()),
                                                         float(beta1), float(beta2), float(eps), ctypes.c_void_p(hyper_dev.data_ptr()),
                                                         int(lo), int(self.n if hi is None else hi), int(slot), int(grid),
                                                         ctypes.c_void_p(stream.cuda_stream)), "r2l_dp_adam_step_range")

    def close(self):
        from . import _lib
        self.grads = self.params = None
        _lib.lib().r2l_dp_destroy()
