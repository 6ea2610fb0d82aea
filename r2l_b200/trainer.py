"""The body of the reference's training loop for the R2L network (main.py:1176-1425) as one object, device-resident and
(on one GPU) replayed as a CUDA graph.

    reference, per iteration                                           here
    ---------------------------------------------------------------   ------------------------------------------------
    LR schedule, param_group['lr'] = ...            (:1181-1195)       r2l_adam_schedule_dev: device-side step counters
    hard-ray pool: np.random.permutation + cat      (:1325-1347)       r2l_pool_draw (one kernel, inside the graph)
    sample_train -> positional_embedder -> model    (:1369-1374)       r2l_forward_train (one kernel)
    img2mse * lw_rgb, psnr.item()                   (:1377-1379)       r2l_mse_loss_grad (one kernel, no host sync)
    optimizer.zero_grad(); loss.backward()          (:1403-1404)       r2l_backward (chain + weight gradients)
    [DataParallel reduce to GPU 0]                  (:472-479)         chunked, overlapped all-reduce of ONE flat buffer (N > 1)
    optimizer.step()  (Adam over 176 tensors)       (:1406)            r2l_adam_step_dev + r2l_pack_weights
    torch.sort of per-ray errors, pool update       (:1410-1425)       r2l_pool_update (one kernel, inside the graph)

No arithmetic of the network happens in torch; torch provides buffers, streams, the CUDA-graph capture and NCCL.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import ops
from .nerf_raybased import N_SAMPLES, NUM_PARAMS


def lr_at(global_step: int, lrate: float, lrate_decay: int, warmup_lr: str | None = None) -> float:
    """Learning rate of iteration `global_step` (main.py:1181-1195; args.lrate, args.lrate_decay in 1000 steps,
    args.warmup_lr = 'start_lr,end_iter')."""
    decay_rate, decay_steps = 0.1, lrate_decay * 1000
    if warmup_lr:
        start_lr, end_iter = [float(x) for x in warmup_lr.split(',')]
        if global_step < end_iter:
            return (lrate - start_lr) / end_iter * global_step + start_lr
        return lrate * (decay_rate ** ((global_step - end_iter) / decay_steps))
    return lrate * (decay_rate ** (global_step / decay_steps))


class HardRayPool:
    """Hard-example pool of main.py:1325-1347 (draw) and :1410-1425 (update), kept on the device and maintained by two
    kernels of the library (csrc/pool.cu) that are part of the captured iteration: no ATen launch, no host sync.

    The reference sorts the per-ray errors and draws replacement slots with np.random.permutation on the host every step.
    Semantics kept: the n_hard_in rays with the largest mean squared error of the FRESH part of the batch enter; until the
    pool holds batch_size * hard_mul rays they are appended, afterwards they overwrite the first n_hard_in of the n_hard_out
    slots drawn this step.  Deliberate differences: the slots are the first n_hard_out values of a keyed pseudo-random
    permutation evaluated on the device (same distribution, another stream of numbers than numpy's), and rays with EQUAL
    error at the selection threshold enter lowest index first (torch.sort leaves that order unspecified).

    The host keeps only a mirror of the fill state (`size`, `full`): it follows from the call sequence alone, so deciding
    the batch size of the next iteration never reads device memory."""

    def __init__(self, batch_size: int, hard_ratio, hard_mul: float, device, seed: int = 0):
        if isinstance(hard_ratio, (list, tuple)):
            n_in, n_out = int(hard_ratio[0] * batch_size), int(hard_ratio[1] * batch_size)
        else:
            n_in = n_out = int(hard_ratio * batch_size)
        self.n_hard_in, self.n_hard_out = min(n_in, n_out), n_out
        if self.n_hard_in <= 0:
            # the reference's indices[-0:] would append the WHOLE batch every step (main.py:1414 with n_hard_in == 0): a quirk
            # nobody relies on; refuse instead of silently doing something else
            raise ValueError(f"HardRayPool: hard_ratio {hard_ratio} selects no ray of a batch of {batch_size}")
        self.hard_mul = hard_mul
        self.fill_level = batch_size * hard_mul      # compared as a float, like main.py:1424
        self.batch_size = batch_size
        # appended in steps of n_hard_in until size >= batch_size * hard_mul (main.py:1423-1425)
        step = self.n_hard_in
        slots = -(-int(self.fill_level) // step) * step + step
        full_size = max(1, -(-self.fill_level // step)) * step           # rays in the pool from the moment it counts as full
        if self.n_hard_out > full_size:
            # the reference would silently draw fewer rays than n_hard_out (permutation(len(pool))[:n_hard_out], main.py:1330);
            # batches of a size that depends on the pool are not supported here
            raise ValueError(f"HardRayPool: n_hard_out = {self.n_hard_out} exceeds the {int(full_size)} rays a full pool holds "
                             f"(hard_ratio {hard_ratio}, hard_mul {hard_mul}, batch {batch_size})")
        self.seed = int(seed)
        self.rays = torch.zeros((slots, 9), dtype=torch.float32, device=device)
        self.state = torch.zeros(1, dtype=torch.int32, device=device)            # rays in the pool, advanced by the update kernel
        self.slots_out = torch.zeros(max(self.n_hard_out, 1), dtype=torch.int32, device=device)   # slots of the last draw
        self.size = 0            # host mirror of state[0]
        self.full = False

    def n_extra(self) -> int:
        """Pool rays the next batch carries (0 while the pool is filling)."""
        return self.n_hard_out if self.full else 0

    def enqueue_draw(self, dst_rows, counters):
        """Pool full: n_hard_out pool rows -> dst_rows [n_hard_out, 9] (the tail of the batch buffer); one launch."""
        ops.pool_draw(self.rays, self.state, self.n_hard_out, self.seed, counters, dst_rows, self.slots_out)

    def enqueue_update(self, rays9, per_ray_err, batch_size, full):
        """per_ray_err: mean squared error per ray of the whole batch; only the fresh rays [:batch_size] compete (batch_size of
        THIS call, main.py:1324,:1411-1413; n_hard_in / n_hard_out stay those of the first batch - the ray-shard loader's
        batches all have N_rand * 4096 rays).  One launch; the host mirror is advanced separately (advance)."""
        ops.pool_update(rays9, per_ray_err, batch_size, self.n_hard_in, self.rays, self.state, self.slots_out if full else None)

    def advance(self, batch_size=None):
        """Host mirror of one update (main.py:1419-1425): call once per iteration whose update was enqueued or replayed."""
        batch_size = self.batch_size if batch_size is None else int(batch_size)
        if not self.full:
            self.size += self.n_hard_in
            if self.size >= batch_size * self.hard_mul:
                self.full = True


class R2LTrainer:
    """One training iteration of R2L per call to step(); state (parameters, Adam moments, hard pool, iteration counter)
    lives on the device.  `model` is a r2l_b200 NeRF_v3_2 on CUDA, `point_sampler` its PointSampler."""

    def __init__(self, model, point_sampler, lrate=5e-4, lrate_decay=500, warmup_lr=None, lw_rgb=1.0, perturb=0.0,
                 hard_ratio=0, hard_mul=1, betas=(0.9, 0.999), eps=1e-8, use_graph=True, group=None, start_step=0,
                 grad_split_layers=(64, 43, 21), comm_sms=0, dp_mode=None, pool_seed=0):
        if not model.flat.is_cuda:
            raise RuntimeError("R2LTrainer: the model must live on a CUDA device (no CPU fallback)")
        self.model, self.sampler = model, point_sampler
        self.lrate, self.lrate_decay, self.warmup_lr, self.lw_rgb = lrate, lrate_decay, warmup_lr, float(lw_rgb)
        self.perturb, self.betas, self.eps = float(perturb), betas, eps
        self.hard_ratio, self.hard_mul, self.pool_seed = hard_ratio, hard_mul, int(pool_seed)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        # N > 1: "peer" = the fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory (csrc/dp.cu, the default
        # where the GPUs of the group can map each other's memory), "nccl" = chunked, overlapped NCCL all-reduces + local Adam
        self.dp_mode = None
        if self.world > 1:
            self.dp_mode = dp_mode or os.environ.get("R2L_DP_MODE") or "peer"
            if self.dp_mode not in ("peer", "nccl"):
                raise ValueError(f"R2LTrainer: dp_mode must be 'peer' or 'nccl', got {self.dp_mode!r}")
        if self.world > 1 and self.dp_mode == "nccl" and group is None and dist.get_backend() == "nccl":
            # own communicator whose NCCL stream has HIGH priority: its kernels are dispatched ahead of the pending CTAs of
            # the weight-gradient launches whenever SMs free up, so the chunk all-reduces really overlap the backward
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
            group = dist.new_group(backend="nccl", pg_options=opts)
        self.group = group
        self.use_graph = bool(use_graph)
        # data parallel: the flat gradient is completed and all-reduced in chunks, top of the buffer first (gradients complete
        # tail -> head), on a communication stream while the backward still runs; only the last chunk's reduction is exposed
        # both data-parallel modes complete the gradient buffer in chunks (top first: gradients complete tail -> head) and
        # exchange a chunk on a HIGH-PRIORITY communication stream as soon as it is complete, while the backward still runs;
        # only the last chunk's exchange is exposed
        dp = self.world > 1
        self.split_layers = list(grad_split_layers) if dp else []
        self.reserve_sms = int(comm_sms) if dp else 0
        self.chunk_ranges = ops.grad_chunk_ranges(self.split_layers) if dp else []
        self.comm_stream = torch.cuda.Stream(model.flat.device, priority=-1) if dp else None
        self.comm_grid = int(os.environ.get("R2L_DP_CHUNK_GRID", "64"))    # CTAs of an overlapped chunk exchange (peer mode)
        self.global_step = start_step          # iterations done; the next one is global_step + 1 (main.py:1175 `start + 1`)
        dev = self.dev = model.flat.device
        self.z_vals = point_sampler.z_vals.tolist()
        lower, diff = point_sampler.jitter_bounds()
        self.z_lower, self.z_diff = lower.tolist(), diff.tolist()
        self.exp_avg = torch.zeros(NUM_PARAMS, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(NUM_PARAMS, dtype=torch.float32, device=dev)
        self.adam_steps = 0
        self.peer = None
        if self.dp_mode == "peer":
            from .parallel import PeerDataParallel
            self.peer = PeerDataParallel(NUM_PARAMS, dev, group)
            with torch.no_grad():      # the parameters move into this rank's peer-visible buffer; the module keeps using them there
                self.peer.params.copy_(model.flat.detach())
                model.flat.data = self.peer.params
            self.grads = self.peer.grads
        else:
            self.grads = torch.empty(NUM_PARAMS, dtype=torch.float32, device=dev)
        self.packed = ops.pack_weights(model.flat.detach())
        self._flat_seen = (model.flat.data_ptr(), model.flat._version)    # the parameters self.packed was built from
        # iteration counters [global_step, adam_steps] and the step scalars derived from them live on the DEVICE
        # (ops.adam_schedule_dev): a host running ahead of the GPU cannot hand an iteration another iteration's learning rate
        self.d_steps = torch.tensor([start_step, 0], dtype=torch.int64, device=dev)
        self.d_hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.h_loss = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.pool = None
        self._static = {}     # n_rays -> dict(buffers, graph)
        self.last_lr = None

    # ---- the device work of one iteration on static buffers (captured once per batch size and entry point) ----
    def _body(self, st, from_host: bool):
        n = st["in9"].shape[0]
        if from_host:     # host-fed iteration: the batch comes from the pinned staging rows, the loss goes back to the host
            st["in9"][:st["h9"].shape[0]].copy_(st["h9"], non_blocking=True)
        batch, pool = st["batch"], self.pool
        if pool is not None and st["pool_full"]:        # hard rays of this iteration -> rows [batch:] (main.py:1325-1347)
            pool.enqueue_draw(st["in9"][batch:], self.d_steps)
        ops.adam_schedule_dev(self.d_steps, self.d_hyper, self.lrate, self.lrate_decay, self.warmup_lr, self.betas[0], self.betas[1])
        kw = dict(rays9=st["in9"])                      # (o | d | rgb) rows: the kernels read the columns in place
        if st["t_rand"] is not None:
            kw.update(t_rand=st["t_rand"], z_lower=self.z_lower, z_diff=self.z_diff)
        else:
            kw.update(z_vals=self.z_vals)
        rgb, ctx = ops.forward_train(self.packed, fwd_saved=st["fwd_saved"], workspace=st["workspace"], **kw)
        n_global = n * self.world
        ops.mse_loss_grad(rgb, st["in9"][:, 6:9], 2.0 * self.lw_rgb / (3 * n_global), self.lw_rgb / (3 * n), grad_rgb=st["grad_rgb"],
                          per_ray_err=st["err"], loss=self.loss)
        if pool is not None:                            # hardest fresh rays -> pool (main.py:1410-1425)
            pool.enqueue_update(st["in9"], st["err"], batch, st["pool_full"])
        ops.backward(self.packed, ctx, st["grad_rgb"], self.grads, bwd_saved=st["bwd_saved"], workspace=st["workspace"],
                     split_layers=self.split_layers, reserve_sms=self.reserve_sms)
        flat = self.model.flat.data
        if self.peer is not None:
            # gradients of all ranks -> my slice -> Adam -> new parameters into every rank's buffer: one kernel per chunk over
            # NVLink peer memory, no NCCL; the chunks that complete early are exchanged beside the rest of the backward
            main = torch.cuda.current_stream(self.dev)
            args = (self.exp_avg, self.exp_avg_sq, self.betas[0], self.betas[1], self.eps, self.d_hyper)
            for i, (lo, hi) in enumerate(self.chunk_ranges[:-1]):
                ops.stream_wait_grad_chunk(i, self.comm_stream)
                self.peer.adam_step(*args, lo=lo, hi=hi, slot=i, grid=self.comm_grid, stream=self.comm_stream)
            lo, hi = self.chunk_ranges[-1]
            self.peer.adam_step(*args, lo=lo, hi=hi, slot=len(self.chunk_ranges) - 1, stream=main)
            if len(self.chunk_ranges) > 1:
                main.wait_stream(self.comm_stream)
            ops.pack_weights(flat, out=self.packed)
            if from_host:
                self.h_loss.copy_(self.loss, non_blocking=True)
            return
        if self.world > 1:
            # ONE flat 23.7 MB gradient buffer, summed over ranks piece by piece as the pieces complete (same collectives in
            # the same order on every rank; NCCL runs them on its own stream in issue order)
            main = torch.cuda.current_stream(self.dev)
            for i, (lo, hi) in enumerate(self.chunk_ranges[:-1]):
                ops.stream_wait_grad_chunk(i, self.comm_stream)
                with torch.cuda.stream(self.comm_stream):
                    dist.all_reduce(self.grads[lo:hi], group=self.group)
            lo, hi = self.chunk_ranges[-1]
            dist.all_reduce(self.grads[lo:hi], group=self.group)     # the head's piece: complete only when the backward is
            if len(self.chunk_ranges) > 1:
                main.wait_stream(self.comm_stream)
        ops.adam_step_dev(flat, self.grads, self.exp_avg, self.exp_avg_sq, self.betas[0], self.betas[1], self.eps, self.d_hyper)
        ops.pack_weights(flat, out=self.packed)                 # operands of the next forward (training or rendering)
        if from_host:
            self.h_loss.copy_(self.loss, non_blocking=True)

    MAX_BATCH_SIZES = 4    # static buffer sets (and graphs) kept; a 4096-ray set is ~1 GB of saved operand images

    def _static_for(self, n, batch=None, pool_full=False):
        batch = n if batch is None else batch
        st = self._static.get(n)
        if st is not None:
            self._static[n] = self._static.pop(n)        # most recently used last
            if (st["batch"], st["pool_full"]) != (batch, pool_full):   # same total, another fresh / pool split: re-capture
                st["graph"] = st["graph_host"] = None
                st["batch"], st["pool_full"] = batch, pool_full
        if st is None:
            while len(self._static) >= self.MAX_BATCH_SIZES:   # e.g. ragged last batches: drop the least recently used set
                old = self._static.pop(next(iter(self._static)))
                old["graph"] = old["graph_host"] = None
                del old
            dev = self.dev
            nf, nb, nw = ops.train_buffer_bytes(n)
            st = dict(in9=torch.zeros((n, 9), device=dev), h9=None,
                      fwd_saved=torch.empty(nf, dtype=torch.uint8, device=dev), bwd_saved=torch.empty(nb, dtype=torch.uint8, device=dev),
                      workspace=torch.empty(nw, dtype=torch.uint8, device=dev),
                      t_rand=torch.zeros((n, N_SAMPLES), device=dev) if self.perturb > 0 else None,
                      grad_rgb=torch.empty((n, 3), device=dev), err=torch.empty(n, device=dev), graph=None, graph_host=None, warm=0,
                      batch=batch, pool_full=pool_full)
            self._static[n] = st
        return st

    def _run(self, st, from_host=False):
        """Run the iteration body on the static buffers: eagerly twice (allocations, lazy CUDA init), then as a graph."""
        if not self.use_graph:
            return self._body(st, from_host)
        key = "graph_host" if from_host else "graph"
        if st[key] is None:
            if st["warm"] < 2:
                st["warm"] += 1
                return self._body(st, from_host)
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.dev)
            with torch.cuda.graph(g):
                self._body(st, from_host)
            st[key] = g
        st[key].replay()

    def _begin(self):
        step = self.global_step + 1
        self.last_lr = lr_at(step, self.lrate, self.lrate_decay, self.warmup_lr)   # host mirror, for logs / state_dict only
        self.adam_steps += 1
        # parameters changed behind the trainer's back (model.load_state_dict after construction, a manual write to
        # model.flat): the packed operand image is stale - rebuild it before this iteration's forward
        flat = self.model.flat
        if (flat.data_ptr(), flat._version) != self._flat_seen:
            ops.pack_weights(flat.detach(), out=self.packed)
            self._flat_seen = (flat.data_ptr(), flat._version)

    def _end(self):
        self.global_step += 1
        # the kernels wrote the parameters and their packed image through raw pointers: tell autograd / the model's cache
        flat = self.model.flat
        torch.autograd.graph.increment_version(flat)
        self._flat_seen = (flat.data_ptr(), flat._version)
        self.model._packed = self.packed
        self.model._packed_version = (flat.data_ptr(), flat._version, str(flat.device))

    def _iterate(self, batch, fill, from_host=False):
        """Common part of the entry points: static buffers of the batch size (fresh rays + the pool rays the iteration will
        draw), `fill(st, batch)` puts the fresh rays into rows [:batch] of the [n, 9] input buffer (or its pinned staging),
        the iteration (pool draw and pool update are launches inside it)."""
        self._begin()
        if self.hard_ratio and self.pool is None:
            self.pool = HardRayPool(batch, self.hard_ratio, self.hard_mul, self.dev, seed=self.pool_seed)
        pool_full = self.pool is not None and self.pool.full
        n = batch + (self.pool.n_extra() if self.pool is not None else 0)
        st = self._static_for(n, batch, pool_full)
        fill(st, batch)
        if st["t_rand"] is not None:
            st["t_rand"].uniform_()           # sample_train's torch.rand (nerf_raybased.py:122), drawn on the device
            given = st.pop("t_rand_given", None)
            if given is not None:             # the caller's uniforms for the fresh rays (pool rays keep device draws)
                st["t_rand"][:batch].copy_(given)
        self._run(st, from_host)
        if self.pool is not None:
            self.pool.advance(batch)
        self._end()
        return self.loss

    @torch.no_grad()
    def step(self, rays_o, rays_d, target, t_rand=None):
        """One iteration on device tensors rays_o, rays_d, target [N,3].  Returns the loss as a 1-element device tensor
        (no host sync; `float(loss)` when the caller wants the number).  perturb > 0 draws t_rand on the device unless given."""
        if t_rand is not None and not self.perturb > 0:
            raise ValueError("R2LTrainer.step: t_rand given but the trainer was built with perturb = 0")

        def fill(st, batch):
            st["in9"][:batch, 0:3].copy_(rays_o, non_blocking=True)
            st["in9"][:batch, 3:6].copy_(rays_d, non_blocking=True)
            st["in9"][:batch, 6:9].copy_(target, non_blocking=True)
            if t_rand is not None:
                st["t_rand_given"] = t_rand
        return self._iterate(rays_o.shape[0], fill)

    @torch.no_grad()
    def step_rays9(self, rays9):
        """One iteration on a DEVICE batch of shard rows [N, 9] = (o | d | rgb), e.g. RayShardLoader.device_batches(packed=True):
        one copy into the static input buffer, the kernels read the columns in place."""
        return self._iterate(rays9.shape[0], lambda st, batch: st["in9"][:batch].copy_(rays9, non_blocking=True))

    @torch.no_grad()
    def step_host(self, rays_o, rays_d=None, target=None):
        """One iteration from HOST tensors: either a [N, 9] batch of shard rows (RayShardLoader.next_buffer()) or the three
        [N, 3] tensors.  The rows are staged in pinned memory; their H2D copy, the iteration and the D2H copy of the loss are
        ONE graph replay, then the host waits for the loss.  Returns the loss as a float."""
        def fill(st, batch):
            if st["h9"] is None or st["h9"].shape[0] != batch:
                st["h9"] = torch.empty((batch, 9), dtype=torch.float32).pin_memory()
                st["graph_host"] = None               # the graph holds the staging address
            if rays_d is None:
                st["h9"].copy_(rays_o)
            else:
                st["h9"][:, 0:3].copy_(rays_o); st["h9"][:, 3:6].copy_(rays_d); st["h9"][:, 6:9].copy_(target)
        self._iterate(rays_o.shape[0], fill, from_host=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return float(self.h_loss[0])

    def close(self):
        """Drop the captured CUDA graphs and static buffers (call before torch.distributed.destroy_process_group(): graphs that
        captured NCCL collectives must be released while the communicator is alive)."""
        torch.cuda.synchronize(self.dev)
        for st in self._static.values():
            st["graph"] = st["graph_host"] = None
        self._static.clear()
        torch.cuda.synchronize(self.dev)
        if self.peer is not None:
            # the module's parameters must not point into the buffer that is about to be freed
            with torch.no_grad():
                self.model.flat.data = self.model.flat.data.clone()
            self.grads = None
            self.peer.close()
            self.peer = None

    def _full_moments(self):
        """Adam moments as full-size tensors: in the peer data-parallel mode every rank holds only its slice (collective call)."""
        if self.peer is None:
            return self.exp_avg, self.exp_avg_sq
        out = []
        for t in (self.exp_avg, self.exp_avg_sq):
            full = torch.zeros_like(t)
            for lo, hi in self.chunk_ranges:
                a, b = self.peer.slice_of(lo, hi)
                full[a:b] = t[a:b]
            dist.all_reduce(full, group=self.group)
            out.append(full)
        return out

    def state_dict(self):
        """Optimizer-side state in torch.optim.Adam's layout for the single flat parameter (ckpt['optimizer_state_dict']).
        (Collective in the peer data-parallel mode: the moments are gathered from the ranks' slices.)"""
        exp_avg, exp_avg_sq = self._full_moments()
        return {"state": {0: {"step": self.adam_steps, "exp_avg": exp_avg, "exp_avg_sq": exp_avg_sq}},
                "param_groups": [{"lr": self.last_lr if self.last_lr is not None else self.lrate, "betas": self.betas,
                                  "eps": self.eps, "params": [0]}], "global_step": self.global_step}

    def load_state_dict(self, sd):
        s = sd["state"][0]
        self.adam_steps = int(s["step"])
        self.exp_avg.copy_(s["exp_avg"]); self.exp_avg_sq.copy_(s["exp_avg_sq"])
        self.global_step = int(sd.get("global_step", self.global_step))
        self.d_steps.copy_(torch.tensor([self.global_step, self.adam_steps], dtype=torch.int64))
