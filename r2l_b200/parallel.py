"""Data-parallel plumbing: one process per GPU, rays sharded, ONE all-reduce of the flat gradient per step.

Replaces the reference's single-process nn.DataParallel (main.py:37-42, :472-479), which re-broadcasts all 23.7 MB of
parameters every forward and reduces gradients to GPU 0.  Rays are independent, so the forward needs no collective; the
backward needs exactly one sum over ranks of the 5,917,187-float gradient buffer.
"""
from __future__ import annotations

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
