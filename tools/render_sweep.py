"""BASELINE config 5: 800x800 inference sweep, ray batches {1k, 4k, 16k, 64k}, 1 -> N GPUs (weak scaling), rays/s against
the tensor roofline; plus render_test split over the ranks with the gathered frames checked bit-equal to a 1-GPU render.

    python tools/render_sweep.py                                            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/render_sweep.py

Reference: main.py:300-309 (render_path: one pose -> sample_test -> embedder -> model, chunked by --chunk), :473 ("when
rendering, use just one GPU").  Rays are independent: every rank renders its own rays, the only communication is the
final gather of the frames (r2l_b200/parallel.py: render_poses_sharded).  Timing: CUDA events on the launching stream, after
warm-up, max over ranks; whole-job rays/s = all ranks' rays / that time.  Prints one JSON line per point (rank 0)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import nerf_raybased as nb  # noqa: E402
from r2l_b200 import ops, parallel  # noqa: E402

FLOP_PER_RAY = 11_789_824


@torch.no_grad()
def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nb.device = dev
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # a frame is a long back-to-back run of the chain kernel: the SUSTAINED cuBLAS bf16 figure is the denominator
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 0)) or 0) or None
    model = nb.NeRF_v3_2(nb.readme_args(), 1008, 3).to(dev)
    with torch.no_grad():
        model.flat.copy_(nb.init_flat_params(0).to(dev))
    H = W = 800
    focal = 555.5555155968841 * 2
    ps = nb.PointSampler(H, W, focal, 16, 2.0, 6.0)
    packed = model.packed_weights()
    z = ps.z_vals.tolist()
    g = torch.Generator(device="cpu").manual_seed(1234)
    poses_all = torch.randn(8, 3, 4, generator=g) * 0.5
    poses_all[:, :, 3] = torch.tensor([0., 0., 4.])

    def agg_ms(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def timed(fn, reps):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return agg_ms(e0.elapsed_time(e1) / reps)

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    # ---- A. batch sweep: one 800x800 frame per GPU (weak scaling), rays fed to the kernel `batch` at a time like --chunk ----
    rays_o, rays_d = ps._pose_rays(poses_all[rank % 8].to(dev))
    rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
    n = rays_o.shape[0]
    out = torch.empty((n, 3), device=dev)
    for batch in (1024, 4096, 16384, 65536, n):
        def frame():
            for lo in range(0, n, batch):
                hi = min(n, lo + batch)
                ops.forward(packed, rays_o=rays_o[lo:hi], rays_d=rays_d[lo:hi], z_vals=z, out=out[lo:hi])
        ms = timed(frame, reps=3 if batch < 16384 else 5)
        rate = world * n / ms * 1e3
        tf = rate * FLOP_PER_RAY / 1e12
        emit(part="batch_sweep", n_gpus=world, frame="800x800", batch=batch, ms_per_frame=round(ms, 4), rays_per_s=round(rate),
             per_gpu_rays_per_s=round(rate / world), tflops_algorithmic=round(tf, 1),
             frac_tensor_peak_algorithmic=(round(tf / world / peak_tf, 4) if peak_tf else None),
             frac_tensor_peak_issued=(round(3 * tf / world / peak_tf, 4) if peak_tf else None))

    # ---- B. pose -> frame kernel, 2 frames per GPU (weak scaling) ----
    my_poses = poses_all[[(2 * rank) % 8, (2 * rank + 1) % 8]].to(dev)
    ms = timed(lambda: model.render_poses(my_poses, ps, focal), reps=5)
    rate = world * 2 * n / ms * 1e3
    emit(part="render_poses", n_gpus=world, frame="800x800", poses_per_gpu=2, ms=round(ms, 4), rays_per_s=round(rate),
         frames_per_s=round(world * 2 / ms * 1e3, 2), tflops_algorithmic=round(rate * FLOP_PER_RAY / 1e12, 1))

    # ---- C. render_test split over the ranks: gathered frames == the 1-GPU frames, bit for bit ----
    for n_poses in (1, 8):
        c2w = poses_all[:n_poses].to(dev)
        frames = parallel.render_poses_sharded(model, c2w, ps, focal)
        ms_sharded = timed(lambda: parallel.render_poses_sharded(model, c2w, ps, focal), reps=3)
        if n_poses >= world or world == 1:
            alone = model.render_poses(c2w, ps, focal)
        else:   # the ray-sharded path goes through the rays -> rgb kernel: compare with the same kernel on all rays
            alone = torch.stack([model.forward_rays(*[t.contiguous() for t in ps._pose_rays(c)], ps).reshape(H, W, 3) for c in c2w])
        same = bool(torch.equal(frames, alone))
        flag = torch.tensor([int(same)], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        emit(part="sharded_render_test", n_gpus=world, poses=n_poses, split="poses" if n_poses >= world else "rays",
             ms_incl_gather=round(ms_sharded, 3), rays_per_s=round(n_poses * n / ms_sharded * 1e3),
             gathered_equals_single_gpu_bitwise=bool(int(flag)))
        if not int(flag):
            raise SystemExit("sharded frames differ from the single-GPU frames")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
