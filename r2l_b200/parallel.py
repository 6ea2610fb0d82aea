"""Data-parallel plumbing: one process per GPU, rays sharded; training = ONE exchange of the flat gradient per step,
rendering = no collective but the final gather of the frames (BASELINE config 5).

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


def gather_rows(local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """Inference: every rank holds rows [shard_range(n, rank, world)) of an [n, ...] result; returns the whole result on
    every rank (rank-order concatenation).  One all_gather of equal-size (padded) pieces; with a gloo group CUDA rows are
    staged through the host."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (shard_range(n, r, world) for r in range(world))]
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"gather_rows: this rank holds {local.shape[0]} rows, its shard of {n} has {sizes[dist.get_rank(group)]}")
    via_host = local.is_cuda and dist.get_backend(group) == "gloo"
    piece = torch.zeros((max(sizes),) + tuple(local.shape[1:]), dtype=local.dtype, device="cpu" if via_host else local.device)
    piece[:local.shape[0]].copy_(local)
    parts = [torch.empty_like(piece) for _ in range(world)]
    dist.all_gather(parts, piece, group=group)
    out = torch.cat([p[:k] for p, k in zip(parts, sizes)], 0)
    return out.to(local.device) if via_host else out


gather_rgb = gather_rows     # name used by the round-1 tests


def render_poses_shard(model, c2w: torch.Tensor, point_sampler, focal: float, rank: int, world: int) -> torch.Tensor:
    """Rank `rank`'s share of rendering the poses c2w[P,3,4] with `world` ranks, as rows [k, 3] of the [P*H*W, 3] result:

      * P >= world: contiguous pose ranges, whole frames through the pose -> frame kernel (r2l_render_poses);
      * P <  world: (one frame at a time is latency-critical) the P*H*W rays are cut into contiguous ranges and run through the
        rays -> rgb kernel on PointSampler's own rays_o / rays_d.

    Rays are independent (main.py:300-309), so no collective is needed to render; gather_rows assembles the frames."""
    if c2w.dim() == 2:
        c2w = c2w[None]
    n_poses, hw = c2w.shape[0], point_sampler.H * point_sampler.W
    if n_poses >= world:
        lo, hi = shard_range(n_poses, rank, world)
        if hi == lo:
            return torch.empty((0, 3), dtype=torch.float32, device=model.flat.device)
        return model.render_poses(c2w[lo:hi], point_sampler, focal).reshape(-1, 3)
    lo, hi = shard_range(n_poses * hw, rank, world)
    parts = []
    for k in range(lo // hw, (hi - 1) // hw + 1 if hi > lo else lo // hw):
        a, b = max(lo, k * hw) - k * hw, min(hi, (k + 1) * hw) - k * hw
        rays_o, rays_d = point_sampler._pose_rays(c2w[k].to(model.flat.device, torch.float32))
        with torch.no_grad():
            parts.append(model.forward_rays(rays_o[a:b].contiguous(), rays_d[a:b].contiguous(), point_sampler))
    return torch.cat(parts, 0) if parts else torch.empty((0, 3), dtype=torch.float32, device=model.flat.device)


def render_poses_sharded(model, c2w: torch.Tensor, point_sampler, focal: float, group=None) -> torch.Tensor:
    """render_test on all GPUs of the job (the reference renders on ONE GPU: main.py:473 "when rendering, use just one GPU"):
    every rank renders its share (render_poses_shard) and all ranks get the frames [P,H,W,3]."""
    on = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
    single = c2w.dim() == 2
    n_poses = 1 if single else c2w.shape[0]
    hw = point_sampler.H * point_sampler.W
    local = render_poses_shard(model, c2w, point_sampler, focal, rank, world)
    if n_poses >= world:      # whole frames per rank: gather in units of frames so that shard sizes follow the pose ranges
        frames = gather_rows(local.reshape(-1, hw, 3), n_poses, group)
    else:
        frames = gather_rows(local, n_poses * hw, group)
    frames = frames.reshape(n_poses, point_sampler.H, point_sampler.W, 3)
    return frames[0] if single else frames


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
