"""world_size-2 gloo test (CPU) of the data-parallel host logic: ray sharding + one flat-gradient all-reduce reproduce the
full-batch gradient.  The per-shard arithmetic is done by the numpy oracle here (the product has no CPU path); on GPUs
the same plumbing runs over NCCL (bench.py, torchrun)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from r2l_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import r2l_oracle as orc
    from r2l_b200.nerf_raybased import init_flat_params
    torch.set_num_threads(2)
    flat = init_flat_params(0).numpy().astype(np.float64)
    rng = np.random.RandomState(0)
    x = rng.randn(n, orc.IN_DIM) * 0.5
    target = rng.rand(n, 3)
    lo, hi = parallel.shard_range(n, rank, world)
    # local gradient of the GLOBAL mean: scale the local-mean gradient by n_local / n_global
    _, g_local, rgb_local, _ = orc.r2l_loss_and_grads(flat, x[lo:hi], target[lo:hi])
    g = torch.from_numpy(g_local * ((hi - lo) / n))
    parallel.allreduce_flat_grads(g)
    rgb = parallel.gather_rows(torch.from_numpy(rgb_local), n)
    # frames gathered in units of whole frames (config 5, P >= world): 3 poses over 2 ranks = 2 + 1
    plo, phi = parallel.shard_range(3, rank, world)
    frames = parallel.gather_rows(torch.arange(3 * 4 * 3, dtype=torch.float32).reshape(3, 4, 3)[plo:phi], 3)
    assert torch.equal(frames, torch.arange(3 * 4 * 3, dtype=torch.float32).reshape(3, 4, 3))
    stub, ps = _StubModel(), _stub_sampler()
    for n_poses in (1, 3):
        c2w = torch.from_numpy(np.random.RandomState(n_poses).randn(n_poses, 3, 4).astype(np.float32))
        got = parallel.render_poses_sharded(stub, c2w, ps, 10.0)
        assert torch.equal(got, stub.render_poses(c2w, ps, 10.0))
    if rank == 0:
        _, g_full, rgb_full, _ = orc.r2l_loss_and_grads(flat, x, target)
        np.save(os.path.join(out_dir, "err.npy"), np.array([
            np.linalg.norm(g.numpy() - g_full) / np.linalg.norm(g_full), np.abs(rgb.numpy() - rgb_full).max()]))
        # the scaling helper is the same statement on tensors
        gr = parallel.mse_grad_rgb(torch.from_numpy(rgb_local), torch.from_numpy(target[lo:hi]), n)
        assert torch.allclose(gr, torch.from_numpy((rgb_local - target[lo:hi]) * 2.0 / (3 * n)))
    dist.barrier()
    dist.destroy_process_group()


class _StubModel:
    """Stands in for NeRF_v3_2 in the sharding tests: rgb = a fixed function of the ray, through the same two entry points."""
    flat = torch.zeros(1)

    @staticmethod
    def _shade(o, d):
        return o * 0.25 + d * 2.0           # exact scalings + one rounding: the same bits whatever the slice length

    def render_poses(self, c2w, ps, focal):
        c2w = c2w[None] if c2w.dim() == 2 else c2w
        return torch.stack([self._shade(*ps._pose_rays(c)).reshape(ps.H, ps.W, 3) for c in c2w])

    def forward_rays(self, rays_o, rays_d, ps):
        return self._shade(rays_o, rays_d)


def _stub_sampler():
    from r2l_b200 import nerf_raybased as nb
    nb.device = torch.device("cpu")
    return nb.PointSampler(5, 7, 10.0, 16, 2.0, 6.0)


def test_render_shards_tile_the_frames():
    """Config 5 sharding arithmetic: for every world size the ranks' shares, concatenated in rank order, are the frames -
    whole poses per rank when there are enough poses, contiguous ray ranges of the frames otherwise."""
    stub, ps = _StubModel(), _stub_sampler()
    for n_poses in (1, 2, 5, 8, 11):
        c2w = torch.from_numpy(np.random.RandomState(n_poses).randn(n_poses, 3, 4).astype(np.float32))
        full = stub.render_poses(c2w, ps, 10.0).reshape(-1, 3)
        for world in (1, 2, 3, 4, 8):
            parts = [parallel.render_poses_shard(stub, c2w, ps, 10.0, r, world) for r in range(world)]
            assert torch.equal(torch.cat(parts, 0), full), (n_poses, world)
            if n_poses < world:
                assert max(p.shape[0] for p in parts) - min(p.shape[0] for p in parts) <= 1


def test_shard_range_partitions():
    for n in (0, 1, 7, 4096, 160000):
        for world in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def test_two_rank_gradient_allreduce_matches_full_batch(tmp_path):
    n = 37   # ragged: shards of 19 and 18 rays
    mp.spawn(_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    err = np.load(tmp_path / "err.npy")
    assert err[0] < 1e-12 and err[1] < 1e-12


def _gather_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for n in (7, 3, 2, 12):                     # ragged (3,2,2), one row each, an EMPTY shard on the last rank, even
        full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) * 0.5
        lo, hi = parallel.shard_range(n, rank, world)
        got = parallel.gather_rows(full[lo:hi].clone(), n)
        assert torch.equal(got, full), (n, rank)
    try:
        parallel.gather_rows(torch.zeros(5, 3), 7)      # 5 rows is nobody's shard of 7 rows over 3 ranks
        raise AssertionError("gather_rows accepted a wrong shard size")
    except ValueError:
        pass
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rows_ragged_and_empty_shards_three_ranks():
    """gather_rows (the only communication of sharded rendering): shard sizes that differ by one, a rank with no rows at all,
    and a wrong local size is refused on every rank before any collective is entered."""
    mp.spawn(_gather_worker, args=(3, _free_port()), nprocs=3, join=True)
