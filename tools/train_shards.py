"""BASELINE config 3 in miniature: R2L distillation training from `.npy` ray shards (N_rand shards of 4096 rays per step,
hard_ratio 0.2) with the loader, the trainer and — under torchrun — one gradient all-reduce per step.

    python tools/train_shards.py [--datadir DIR] [--N_rand 2] [--steps 200] [--hard_ratio 0.2] [--hard_mul 20]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_shards.py ...

Without --datadir, synthetic shards are written to a temporary directory (rays of random lego-style poses, targets = a
smooth function of the ray, so the loss visibly falls)."""
import argparse, os, sys, tempfile, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import nerf_raybased as nb
from r2l_b200.data import RayShardLoader, list_shards, write_ray_shards
from r2l_b200.trainer import R2LTrainer


def synthetic_shards(datadir, n_shards, rows, seed=0):
    rng = np.random.RandomState(seed)
    n = n_shards * rows
    d = rng.randn(n, 3).astype(np.float32) * 0.3 + np.array([0, 0, -1], np.float32)
    o = rng.randn(n, 3).astype(np.float32) * 0.1 + np.array([0, 0, 4], np.float32)
    rgb = 1.0 / (1.0 + np.exp(-3.0 * d[:, :3] + o * 0.2))
    return write_ray_shards(np.concatenate([o, d, rgb.astype(np.float32)], 1), datadir, rows, rng=rng)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--datadir", default=None); ap.add_argument("--N_rand", type=int, default=2)
    ap.add_argument("--rows", type=int, default=4096); ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--hard_ratio", type=float, default=0.2); ap.add_argument("--hard_mul", type=float, default=20)
    ap.add_argument("--lrate", type=float, default=5e-4); ap.add_argument("--lrate_decay", type=int, default=500)
    ap.add_argument("--warmup_lr", default="0.0001,200")      # the README's R2L command line
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local); dev = torch.device("cuda", local); nb.device = dev
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    datadir = a.datadir
    if datadir is None:
        datadir = os.path.join(tempfile.gettempdir(), "r2l_synthetic_shards")
        if rank == 0 and not os.path.isdir(datadir):
            synthetic_shards(datadir, max(8 * world, 4 * a.N_rand * world), a.rows)
        if world > 1:
            dist.barrier()
    shards, _, _ = list_shards(datadir, pseudo_ratio=-1)
    torch.manual_seed(0)                       # identical initial weights on every rank
    model = nb.NeRF_v3_2(nb.readme_args(), 1008, 3).to(dev)
    ps = nb.PointSampler(400, 400, 555.5555155968841, 16, 2.0, 6.0)
    trainer = R2LTrainer(model, ps, lrate=a.lrate, lrate_decay=a.lrate_decay, warmup_lr=a.warmup_lr or None, hard_ratio=a.hard_ratio, hard_mul=a.hard_mul)
    loader = RayShardLoader(shards, a.N_rand, rows=a.rows, rank=rank, world=world, seed=1)
    batches = loader.device_batches(dev, packed=True)      # [N, 9] shard rows, read in place by the kernels
    t0, first = None, None
    for it in range(a.steps):
        if it == min(20, a.steps // 2):
            torch.cuda.synchronize(); t0, it0 = time.perf_counter(), it
        loss = trainer.step_rays9(next(batches))
        if it % 50 == 0 or it == a.steps - 1:
            v = float(loss)
            first = v if first is None else first
            if rank == 0:
                print(f"iter {trainer.global_step}: loss {v:.6f} psnr {-10 * np.log10(v):.2f} lr {trainer.last_lr:.3e} pool {'full' if trainer.pool is not None and trainer.pool.full else 'filling'}", flush=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    rays = (a.steps - it0) * a.N_rand * a.rows * world
    if rank == 0:
        print(f"{world} GPU(s): {rays / dt / 1e6:.2f} M fresh rays/s over {a.steps - it0} steps ({1e3 * dt / (a.steps - it0):.3f} ms/step, batch {a.N_rand * a.rows} fresh rays per GPU); loss {first:.5f} -> {float(loss):.5f}")
    loader.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
