#!/usr/bin/env python
"""bench.py — R2L hot-path benchmark (contract: see the task statement / DESIGN.md "Measurement").

Metric (BASELINE.json): rays/sec of the W256/D88 ResMLP light-field network, forward + backward, batch 4096
synthetic rays per GPU, fp32-parity arithmetic (bf16x3 split on tcgen05).  One "step" = one full training pass of
the hot path over one batch: weight re-pack, fused forward (sampling + positional encoding + 88 layers), MSE
gradient, fused backward (chain + weight gradients), gradient all-reduce when N > 1, Adam update.

    python bench.py [--gpus N] [--steps K] [--warmup W]            ours, N = 1
    torchrun ... bench.py --gpus N --steps K --warmup W             ours, N ranks (weak scaling: 4096 rays / GPU)
    python bench.py --impl reference ...                             the reference's CPU code path (stock PyTorch ops)

`value` is measured with inputs resident in HBM; `e2e` goes through the public module API (NeRF_v3_2 +
autograd) from pinned HOST buffers with the H2D copy of the rays/targets and the D2H read of the loss inside the
timed region.  Only the `cpu_baseline` leg and `--impl reference` touch oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 4096
FWD_FLOP_PER_RAY = 11_789_824          # 5,894,912 MAC x 2
BWD_FLOP_PER_RAY = 23_063_552
METRIC = "rays/sec (W256D88 ResMLP fwd+bwd, batch 4096)"
WORKLOAD = ("lego_noview W256D88 ResMLP train step (fwd+bwd+Adam) on 4096 synthetic lego-pose rays per GPU, "
            "random-init weights (seed 0), perturb=0")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops": float(d["bf16_tflops"]), "hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json, burst)"}
    return {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def synthetic_rays(n, seed):
    """Rays of a random lego-style pose (400x400, focal 555.56, radius 4) as in SURVEY.md section 8(d)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    theta, phi = rng.uniform(-180, 180), rng.uniform(-90, 0)
    ct, st, cp, sp = np.cos(np.radians(theta)), np.sin(np.radians(theta)), np.cos(np.radians(phi)), np.sin(np.radians(phi))
    # camera-to-world of pose_spherical(theta, phi, 4) (dataset/load_blender.py:22-28), composed by hand
    t = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 4.0], [0, 0, 0, 1]])
    rp = np.array([[1, 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1]])
    rt = np.array([[ct, 0, -st, 0], [0, 1, 0, 0], [st, 0, ct, 0], [0, 0, 0, 1]])
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    c2w = (flip @ rt @ rp @ t).astype(np.float32)
    pix = rng.choice(400 * 400, size=n, replace=n > 160000)
    i, j = (pix % 400).astype(np.float32), (pix // 400).astype(np.float32)
    focal = np.float32(555.5555155968841)
    dirs = np.stack([(i - 200) / focal, -(j - 200) / focal, -np.ones_like(i)], -1)
    rays_d = (dirs[:, None, :] * c2w[:3, :3]).sum(-1).astype(np.float32)
    rays_o = np.broadcast_to(c2w[:3, 3], rays_d.shape).astype(np.float32).copy()
    target = rng.rand(n, 3).astype(np.float32)
    return rays_o, rays_d, target


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md "clocks" line), read through NVML every
    2 ms (a 20-step timed region is only ~20 ms long: one nvidia-smi process per sample would see it once at best);
    falls back to polling nvidia-smi when NVML cannot be opened.  Samples taken while `active` is set are the ones
    summarised as "under load"."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.active = index, [], False, False
        self.max_mhz, self.source = None, "nvml"

    def _open_nvml(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        return pynvml, h

    def run(self):
        try:
            nv, h = self._open_nvml()
        except Exception:
            return self._run_smi()
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append((self.active, float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons_fn(h))))
            except Exception:
                pass
            time.sleep(0.002)

    def _run_smi(self):
        self.source = "nvidia-smi"
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                if len(f) >= 6:
                    self.max_mhz = float(f[1])
                    mask = sum(b for b, v in zip(bits, f[2:6]) if v.lower().startswith("active"))
                    self.samples.append((self.active, float(f[0]), mask))
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        import statistics
        load = [s for s in self.samples if s[0]] or self.samples
        if not load:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "source": self.source}
        mask = 0
        for s in load:
            mask |= s[2]
        return {"sm_mhz": statistics.median(s[1] for s in load), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "samples": len(load), "source": self.source}


# ------------------------------------------------------------------------------------------------
# the reference's CPU code path (stock PyTorch ops on host cores) — oracle/, checker & baseline only
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(n_rays, threads=None):
    import torch
    from oracle import r2l_oracle as orc
    from oracle.torch_reference import RefR2L, embed, sample
    from r2l_b200.nerf_raybased import init_flat_params
    torch.set_num_threads(threads or os.cpu_count() or 1)
    torch.autograd.set_detect_anomaly(False)
    model = RefR2L().load_flat(init_flat_params(0))
    ro, rd, tg = synthetic_rays(n_rays, 0)
    ro, rd, tg = torch.from_numpy(ro), torch.from_numpy(rd), torch.from_numpy(tg)
    z = torch.from_numpy(orc.sampler_z_vals(2.0, 6.0))
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)

    def step():
        opt.zero_grad()
        loss = ((model(embed(sample(ro, rd, z))) - tg) ** 2).mean()
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def best_cpu_threads(n_rays):
    """The stock PyTorch CPU path does not scale to every core of a big host on these small GEMMs; give the
    reference its best configuration: try a few intra-op thread counts on one step each and keep the fastest."""
    import torch
    cores = os.cpu_count() or 1
    best = (None, float("inf"))
    for th in sorted({min(cores, c) for c in (8, 16, 32, 64, cores)}):
        step = cpu_reference_step_fn(n_rays, th)
        step()
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (th, dt)
        elif dt > 1.5 * best[1]:
            break   # more threads are only getting slower
    return best[0]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    sample_rays = 1024   # bounded sample of the 4096-ray batch per step (same per-ray work)
    threads = best_cpu_threads(sample_rays)
    step = cpu_reference_step_fn(sample_rays, threads)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    rays_s = sample_rays * args.steps / dt
    cores = threads
    line = {"impl": "reference", "metric": METRIC, "value": rays_s, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * (BATCH / sample_rays), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference code path = stock PyTorch CPU ops (oracle/torch_reference.py restates model/nerf_raybased.py; /root/reference is absent on the GPU box)"},
            "cpu_baseline": {"value": rays_s, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} train steps (fwd+bwd+Adam) on {sample_rays} of the 4096 rays, torch {torch.__version__} CPU, {cores} threads (best of 8/16/32/64/all on this {os.cpu_count()}-core host)"},
            "e2e": {"value": rays_s, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from r2l_b200 import ops
    from r2l_b200.nerf_raybased import NeRF_v3_2, PointSampler, init_flat_params, readme_args

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    if os.environ.get("R2L_PAIR_MODE") is not None:   # experiment switch: 0 / 1 / -1 (default)
        from r2l_b200 import _lib as _l
        _l.lib().r2l_set_pair_mode(int(os.environ["R2L_PAIR_MODE"]))

    ro, rd, tg = synthetic_rays(BATCH, seed=rank)
    h_ro, h_rd, h_tg = (torch.from_numpy(a).pin_memory() for a in (ro, rd, tg))
    d_ro, d_rd, d_tg = h_ro.to(dev), h_rd.to(dev), h_tg.to(dev)
    z_vals = PointSampler(400, 400, 555.5555155968841, 16, 2.0, 6.0).z_vals.tolist()

    # ---- device-resident arm: the trainer's iteration (r2l_b200.trainer.R2LTrainer) on rays already in HBM ----
    from r2l_b200.trainer import R2LTrainer
    model = NeRF_v3_2(readme_args(), 1008, 3).to(dev)
    with torch.no_grad():
        model.flat.copy_(init_flat_params(0).to(dev))
    ps = PointSampler(400, 400, 555.5555155968841, 16, 2.0, 6.0)
    trainer = R2LTrainer(model, ps, lrate=5e-4, lrate_decay=500)     # CUDA graph on one GPU, eager + all-reduce on N > 1
    packed = trainer.packed
    n_global = BATCH * world
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    d9 = torch.cat([d_ro, d_rd, d_tg], dim=1).contiguous()      # a device batch as the ray-shard loader yields it: [N, 9] rows

    def step_device():
        trainer.step_rays9(d9)
    # forward chain, loss+grad, backward chain, tail gradients, weight gradients, Adam, pack
    LAUNCHES_PER_STEP = 7

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.active = True
    for a, b in evs:
        flush.fill_(1)          # evict L2 between timed iterations (not timed)
        a.record()
        step_device()
        b.record()
    barrier()
    sampler.active = False
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n_global * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel alone: the forward chain kernel (same kernel template drives the backward chain) ----
    reps = 10
    for _ in range(2):
        ops.forward_train(packed, rays_o=d_ro, rays_d=d_rd, z_vals=z_vals)
    torch.cuda.synchronize()
    k_ms = 0.0
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rgb, ctx = ops.forward_train(packed, rays_o=d_ro, rays_d=d_rd, z_vals=z_vals)
        b.record()
        torch.cuda.synchronize()
        k_ms += a.elapsed_time(b)
        del ctx
    k_ms /= reps
    achieved_tflops = BATCH * FWD_FLOP_PER_RAY / (k_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "r2l_chain_kernel<kFwdTrain, half form> (4096 rays = 32 tiles, one CTA pair per tile: 64 CTAs, tcgen05 cta_group::2 M=128)", "achieved": achieved_tflops,
                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved_tflops / peaks["bf16_tflops"],
                "traffic": 359.4e6, "traffic_source": "profiles/r1_summary.md section 2c: dram__bytes_read.sum + dram__bytes_write.sum of this kernel at 4096 rays (ncu --set full, profiles/r1_full_4096_raw.csv)", "peak_source": peaks["source"], "kernel_ms": k_ms,
                "note": "algorithmic fp32 FLOPs; the kernel issues 3x that as bf16 MMAs (hi*hi+lo*hi+hi*lo) to meet the 1e-3 fp32 parity bar, and a 4096-ray batch (32 tiles of 128 rays, two SMs per tile) occupies 64 of 148 SMs"}

    # ---- the same kernel with every SM busy (148 tiles = 18,944 rays), inference form: kernel quality, not the metric ----
    n_full = 148 * 128
    fo, fd, _ = synthetic_rays(n_full, seed=7)
    fo, fd = torch.from_numpy(fo).to(dev), torch.from_numpy(fd).to(dev)
    fout = torch.empty(n_full, 3, device=dev)
    for _ in range(3):
        ops.forward(packed, rays_o=fo, rays_d=fd, z_vals=z_vals, out=fout)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.forward(packed, rays_o=fo, rays_d=fd, z_vals=z_vals, out=fout)
    b.record()
    torch.cuda.synchronize()
    full_ms = a.elapsed_time(b) / 10
    full_tflops = n_full * FWD_FLOP_PER_RAY / (full_ms * 1e-3) / 1e12
    roofline["full_chip"] = {"kernel": "r2l_chain_kernel<kFwdInfer> (18,944 rays = one tile per SM)", "kernel_ms": full_ms,
                             "achieved": full_tflops, "frac": full_tflops / peaks["bf16_tflops"],
                             "issued_frac": 3 * full_tflops / peaks["bf16_tflops"]}

    # ---- end to end through the public API, host buffers: pinned rays/targets -> device, one iteration, loss -> host ----
    h9 = torch.cat([h_ro, h_rd, h_tg], dim=1).pin_memory()      # a batch as the ray-shard loader yields it: [N, 9] rows

    def step_e2e():
        return trainer.step_host(h9)

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_global * args.steps / float(t.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    line = None
    if rank == 0:
        cpu = None
        if world == 1:
            sample_rays, reps_cpu = 1024, 3
            cpu_threads = best_cpu_threads(sample_rays)
            stepc = cpu_reference_step_fn(sample_rays, cpu_threads)
            stepc()
            t0 = time.perf_counter()
            for _ in range(reps_cpu):
                stepc()
            dt = time.perf_counter() - t0
            cores = cpu_threads
            cpu = {"value": sample_rays * reps_cpu / dt, "unit": "rays/s", "cores": cores, "kind": "port",
                   "sample": f"{reps_cpu} train steps on {sample_rays} of the 4096 rays; stock PyTorch CPU ops (oracle/torch_reference.py), {cores} threads (best of 8/16/32/64/all on this {os.cpu_count()}-core host)"}
        line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (bf16x3 split operands, fp32 accumulate)", "data": "synthetic",
                "config": {"workload": WORKLOAD, "rays_per_gpu": BATCH, "global_batch": n_global,
                           "parallelism": f"dp{world}" if world > 1 else "single",
                           "l2": "256 MiB buffer written between timed iterations (L2 flush, untimed)",
                           "step": "R2LTrainer.step_rays9 on a resident [N,9] batch: forward_train + mse loss/grad + backward(chain, dW, tail) + allreduce(N>1) + Adam + pack_weights"
                                   + (" (one CUDA graph replay)" if trainer.use_graph else " (eager launches)")},
                "clocks": sampler.summary(), "gpu_launches": LAUNCHES_PER_STEP * args.steps,
                "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": BATCH * 9 * 4, "d2h_bytes_per_step": 4,
                        "api": "R2LTrainer.step_host: host batch of [N,9] ray-shard rows -> pinned staging -> device every step, one training iteration, loss read back to the host"},
                "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
