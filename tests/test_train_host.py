"""CPU tests of the host-side training pieces: ray-shard pipeline (row N2), LR schedule and hard-ray pool (row N3)."""
import os

import numpy as np
import pytest
import torch

from oracle import r2l_oracle as orc
from r2l_b200 import data as rd
from r2l_b200.trainer import HardRayPool, lr_at


@pytest.fixture()
def shard_dir(tmp_path):
    rays = np.random.RandomState(0).rand(64 * 10 + 17, 9).astype(np.float32)
    paths = rd.write_ray_shards(rays, str(tmp_path), split_size=64, rng=np.random.RandomState(1))
    return rays, paths, str(tmp_path)


def test_write_ray_shards_format(shard_dir):
    rays, paths, d = shard_dir
    assert len(paths) == 10 and os.path.basename(paths[0]) == "data_1.npy"      # remainder of 17 rays dropped (create_data.py:864)
    got = np.concatenate([np.load(p) for p in paths])
    assert got.dtype == np.float32 and got.shape == (640, 9)
    # a permutation of (a subset of) the input rows
    key = lambda a: {r.tobytes() for r in a}
    assert key(got) <= key(rays) and len(key(got)) == 640


def test_native_reader_matches_np_load(shard_dir):
    _, paths, d = shard_dir
    out = torch.empty(64 * 4, 9)
    rd.read_shards_into(paths[2:6], out, threads=3)
    for k, p in enumerate(paths[2:6]):
        assert np.array_equal(out[64 * k:64 * (k + 1)].numpy(), np.load(p))
    # version-2 header, wrong dtype, wrong size, missing file: loud errors naming the file
    np.save(os.path.join(d, "f64.npy"), np.zeros((64, 9)))
    with pytest.raises(RuntimeError, match="f64.npy"):
        rd.read_shards_into([os.path.join(d, "f64.npy")], torch.empty(64, 9))
    with pytest.raises(RuntimeError, match="expected"):
        rd.read_shards_into([paths[0]], torch.empty(32, 9))
    with pytest.raises(RuntimeError, match="cannot open"):
        rd.read_shards_into([os.path.join(d, "nope.npy")], torch.empty(64, 9))
    with open(os.path.join(d, "v2.npy"), "wb") as f:
        np.lib.format.write_array(f, np.load(paths[0]), version=(2, 0))
    rd.read_shards_into([os.path.join(d, "v2.npy")], out[:64])
    assert np.array_equal(out[:64].numpy(), np.load(paths[0]))
    rd.read_shards_into([], torch.empty(0, 9))


def test_loader_epochs_cover_every_shard_and_match_reference_reader(shard_dir):
    _, paths, d = shard_dir
    ds = rd.BlenderDataset_v2(d, pseudo_ratio=-1)
    assert len(ds) == 10
    by_bytes = {np.load(p).tobytes(): p for p in paths}
    ld = rd.RayShardLoader(paths, shards_per_batch=2, rows=64, seed=5, pin=False, workers=2)
    try:
        seen = []
        for _ in range(10):          # two epochs of 5 batches
            o, dd, t = ld.next()
            assert o.shape == (128, 3) and dd.shape == (128, 3) and t.shape == (128, 3)
            full = torch.cat([o, dd, t], -1).numpy()
            for k in range(2):
                seen.append(by_bytes[full[64 * k:64 * (k + 1)].tobytes()])
        assert sorted(seen[:10]) == sorted(paths) and sorted(seen[10:]) == sorted(paths)
        assert seen[:10] != seen[10:]                     # a fresh permutation per epoch
    finally:
        ld.close()
    # the item format of the reference's dataset class
    o, dd, t = ds[3]
    ref = np.load(ds.all_splits[3])
    assert np.array_equal(o.numpy(), ref[:, :3]) and np.array_equal(dd.numpy(), ref[:, 3:6]) and np.array_equal(t.numpy(), ref[:, 6:9])


def test_loader_rank_partition_is_disjoint(shard_dir):
    _, paths, _ = shard_dir
    parts = []
    for r in range(2):
        ld = rd.RayShardLoader(paths, 1, rows=64, seed=0, rank=r, world=2, pin=False)
        parts.append(set(ld.paths))
        ld.close()
    assert parts[0].isdisjoint(parts[1]) and parts[0] | parts[1] == set(paths)


def test_list_shards_pseudo_ratio(tmp_path):
    for k in range(8):
        np.save(tmp_path / f"data_{k}.npy", np.zeros((4, 9), np.float32))
    for k in range(2):
        np.save(tmp_path / f"train_{k}.npy", np.zeros((4, 9), np.float32))
    allp, n_orig, n_pseudo = rd.list_shards(str(tmp_path), pseudo_ratio=-1)
    assert (len(allp), n_orig, n_pseudo) == (10, 2, 8)
    # pseudo_ratio 0.5: as many pseudo shards as original ones (load_blender.py:285-288)
    allp, _, _ = rd.list_shards(str(tmp_path), pseudo_ratio=0.5, rng=np.random.RandomState(0))
    assert len(allp) == 4 and sum(os.path.basename(p).startswith("train_") for p in allp) == 2


def test_lr_schedule_matches_reference_formula():
    for step in (1, 10, 1999, 2000, 2001, 250000, 1000000):
        assert lr_at(step, 5e-4, 500) == orc.lr_schedule(step, 5e-4, 500)
        assert lr_at(step, 5e-4, 500, "0.0001,2000") == orc.lr_schedule(step, 5e-4, 500, "0.0001,2000")
    assert lr_at(500000, 5e-4, 500) == pytest.approx(5e-5)           # one decade per lrate_decay*1000 steps
    assert lr_at(1000, 5e-4, 500, "0.0001,2000") == pytest.approx(3e-4)


def test_hard_ray_pool_host_mirror_follows_the_reference_fill():
    """The host mirror of the pool (sizes, fullness, rays carried per batch) against the numpy restatement of
    main.py:1410-1425; the device side (selection, slots, row moves) is tests/test_gpu_train.py."""
    rng = np.random.RandomState(0)
    for batch, hard_ratio, hard_mul in ((40, 0.2, 1), (40, 0.2, 2.5), (64, [0.1, 0.3], 1), (50, 0.33, 1.3)):
        pool = HardRayPool(batch, hard_ratio, hard_mul, torch.device("cpu"))
        if isinstance(hard_ratio, list):
            n_in, n_out = int(hard_ratio[0] * batch), int(hard_ratio[1] * batch)
        else:
            n_in = n_out = int(hard_ratio * batch)
        assert (pool.n_hard_in, pool.n_hard_out) == (n_in, n_out)
        ref_rays, ref_full = np.zeros((0, 9), np.float32), False
        for it in range(40):
            assert pool.full == ref_full and pool.n_extra() == (n_out if ref_full else 0)
            n = batch + pool.n_extra()
            o, d, t, rgb = (rng.rand(n, 3).astype(np.float32) for _ in range(4))
            slots = rng.permutation(ref_rays.shape[0])[:n_out] if ref_full else None
            ref_rays, ref_full = orc.hard_pool_update(ref_rays, ref_full, o, d, t, rgb, batch, n_in, hard_mul, slots)
            pool.advance(batch)
            assert pool.full == ref_full and pool.size == ref_rays.shape[0] <= pool.rays.shape[0]
        assert pool.full
    with pytest.raises(ValueError):
        HardRayPool(4, 0.2, 1, torch.device("cpu"))       # int(0.8) == 0 rays: refused (see the class docstring)
    with pytest.raises(ValueError):
        HardRayPool(100, [0.1, 0.3], 0.1, torch.device("cpu"))   # a full pool of 10 rays cannot supply 30 per batch
    assert HardRayPool(100, 0.2, 0.1, torch.device("cpu")).n_hard_out == 20      # full after one append of 20 rays: fine


def test_pool_slot_permutation_is_a_bijection_and_varies_with_the_step():
    """r2l_pool_slot_host = the function r2l_pool_draw evaluates per drawn ray (csrc/pool.cu: pool_slot, same code on host and
    device): for every pool size it is a permutation of [0, size), so the n_hard_out drawn slots are distinct
    (np.random.permutation(n)[:n_hard_out], main.py:1330-1332); it changes with the iteration counter and the seed, and every
    slot is drawn about equally often."""
    from r2l_b200 import ops
    for size in (1, 2, 3, 5, 16, 17, 640, 1000, 4097):
        perm = [ops.pool_slot_host(j, size, 7, 3) for j in range(size)]
        assert sorted(perm) == list(range(size)), size
    a = [ops.pool_slot_host(j, 640, 0, 10) for j in range(128)]
    b = [ops.pool_slot_host(j, 640, 0, 11) for j in range(128)]
    c = [ops.pool_slot_host(j, 640, 1, 10) for j in range(128)]
    assert a != b and a != c and len(set(a)) == 128
    assert 10 < len(set(a) & set(b)) < 50                 # two independent 128-subsets of 640 share 25.6 slots on average
    # first-drawn slot over 4000 iterations of a 100-slot pool: 40 hits per slot expected, sigma 6.3
    hits = np.bincount([ops.pool_slot_host(0, 100, 0, s) for s in range(4000)], minlength=100)
    assert hits.min() >= 15 and hits.max() <= 70, (hits.min(), hits.max())
    chi2 = float(((hits - 40.0) ** 2 / 40.0).sum())
    assert chi2 < 160, chi2                               # 99 degrees of freedom: mean 99, sigma 14
    with pytest.raises(ValueError):
        ops.pool_slot_host(5, 5, 0, 0)


def test_camera_poses_and_rays_match_the_reference(tmp_path):
    """pose_spherical / get_rays of the pseudo-data generator against outputs of the reference's own functions
    (tests/golden/camera_seed0.npz, made by make_golden.py), for the oracle and for the product's host glue."""
    from r2l_b200 import pseudo_data as pd
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_seed0.npz")))
    for (theta, phi), want in zip(g["angles"], g["c2w"]):
        np.testing.assert_allclose(orc.pose_spherical(theta, phi, float(g["radius"])), want, rtol=0, atol=3e-7)
        assert np.array_equal(pd.pose_spherical(float(theta), float(phi), float(g["radius"])).numpy(), want)
    H, W, focal = int(g["H"]), int(g["W"]), float(g["focal"])
    c2w = g["c2w"][3][:3, :4]
    ro, rd = orc.get_rays(H, W, focal, c2w)
    np.testing.assert_allclose(rd, g["rays_d"], rtol=0, atol=2e-7)
    assert np.array_equal(ro, g["rays_o"])
    tro, trd = pd.get_rays(H, W, focal, torch.from_numpy(c2w))
    assert np.array_equal(trd.numpy(), g["rays_d"]) and np.array_equal(tro.numpy(), g["rays_o"])
    # random poses: on the radius-4 sphere, upper hemisphere, reproducible from a seeded generator
    rng = np.random.RandomState(0)
    poses = [pd.get_rand_pose(rng).numpy() for _ in range(20)]
    for p in poses:
        assert abs(np.linalg.norm(p[:3, 3]) - 4.0) < 1e-5 and p[2, 3] >= -1e-6
        np.testing.assert_allclose(p[:3, :3] @ p[:3, :3].T, np.eye(3), atol=1e-6)
    assert np.array_equal(pd.get_rand_pose(np.random.RandomState(0)).numpy(), poses[0])


def _radix_select_like_pool_cu(err: np.ndarray, k: int):
    """numpy restatement of r2l_pool_update_kernel's selection (csrc/pool.cu): four 8-bit radix passes from the top digit find
    the k-th largest key (keys = the float bits; -0.0 folded onto +0.0), then the rays above it in index order followed by
    the lowest-index rays equal to it."""
    keys = err.astype(np.float32).view(np.uint32).copy()
    keys[keys == 0x80000000] = 0
    prefix, mask, remaining = 0, 0, k
    for p in (3, 2, 1, 0):
        shift = 8 * p
        hist = np.bincount(((keys[(keys & mask) == prefix] >> shift) & 255).astype(np.int64), minlength=256)
        b = 255
        while b > 0 and hist[b] < remaining:
            remaining -= hist[b]
            b -= 1
        prefix |= b << shift
        mask |= 255 << shift
    above = np.nonzero(keys > prefix)[0]
    equal = np.nonzero(keys == prefix)[0][:remaining]
    assert len(above) == k - remaining
    return np.concatenate([above, equal])


def test_pool_selection_algorithm_equals_the_reference_sort():
    """The selection rule of the pool update (radix select on the float bits) picks what main.py:1411-1414 picks -
    torch.sort(err).indices[-n_hard_in:] - on distinct errors, and a valid top-k (lowest index first among equals) on ties;
    the CUDA implementation of the same steps is checked on the GPU (tests/test_gpu_train.py)."""
    rng = np.random.RandomState(0)
    for n, k in ((40, 8), (640, 128), (4096, 819), (5000, 1), (1030, 1030)):
        err = (rng.permutation(n).astype(np.float32) / 7.0) ** 2
        got = _radix_select_like_pool_cu(err, k)
        assert sorted(got.tolist()) == sorted(np.argsort(err, kind="stable")[-k:].tolist())
    err = np.array([0.5, 0.25, 0.5, -0.0, 0.5, np.nan, 0.0, 0.5, 0.125], np.float32)
    assert _radix_select_like_pool_cu(err, 3).tolist() == [5, 0, 2] and _radix_select_like_pool_cu(err, 8).tolist() == [0, 1, 2, 4, 5, 7, 8, 3]
    err = rng.randint(0, 5, 1000).astype(np.float32)
    for k in (1, 10, 500, 999, 1000):
        got = _radix_select_like_pool_cu(err, k)
        assert len(set(got.tolist())) == k and err[got].min() >= np.sort(err)[-k]
