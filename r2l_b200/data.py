"""Ray-shard input pipeline: the reference's `[4096, 9]` fp32 `.npy` ray shards, host to device.

Reference: the writer splits shuffled rays (o | d | rgb) into files of `split_size` rows (utils/create_data.py:853-872,
utils/convert_original_data_to_rays_blender.py:215-233); the reader is BlenderDataset_v2 (dataset/load_blender.py:257-324)
behind a torch DataLoader with batch_size = N_rand shards, pin_memory and an infinite random sampler (main.py:795-808), and
the train loop flattens the batch to [N_rand*4096, 3] triples (main.py:1305-1311).  At B200 speed a step consumes a batch
of shards every ~1 ms per 4096 rays, so per-file `np.load` -> `torch.Tensor` -> collate -> pin copies do not keep up.  Here:

* `RayShardLoader` reads shard payloads straight into a ring of PINNED batch buffers through the C library's native
  reader (`r2l_read_ray_shards`: pread on native threads, no intermediate arrays, no GIL), `depth` batches ahead;
* each batch is one `[n_shards * rows, 9]` buffer: rays_o / rays_d / target are column views, and the device copy is ONE
  async H2D of the whole buffer on a side stream, overlapped with the previous step (`device_batches()`);
* under torch.distributed every rank draws from its own slice of the shard list (rank r takes shards r, r + world, ...):
  no rank reads a file another rank uses in the same epoch and no collective is involved.

`BlenderDataset_v2` keeps the reference's class name, constructor arguments and item format for code that indexes it.
"""
from __future__ import annotations

import os
import queue
import threading

import numpy as np
import torch


def write_ray_shards(data, datadir: str, split_size: int = 4096, first_index: int = 1, prefix: str = "data_", shuffle: bool = True, rng=None):
    """rays [N, 9] (o | d | rgb) -> `{datadir}/{prefix}{k}.npy` files of `split_size` rows, shuffled twice as the reference
    does (utils/create_data.py:857-861); the remainder that does not fill a shard is dropped (:864).  Returns the paths."""
    data = np.asarray(data, dtype=np.float32)
    if data.ndim != 2 or data.shape[1] != 9:
        raise ValueError(f"write_ray_shards: expected [N, 9] rays, got {data.shape}")
    os.makedirs(datadir, exist_ok=True)
    if shuffle:
        rng = rng or np.random
        data = data[rng.permutation(data.shape[0])][rng.permutation(data.shape[0])]
    paths = []
    for k, ix in enumerate(range(0, data.shape[0] // split_size * split_size, split_size)):
        path = os.path.join(datadir, f"{prefix}{first_index + k}.npy")
        np.save(path, data[ix:ix + split_size])
        paths.append(path)
    return paths


def list_shards(datadir: str, pseudo_ratio: float = 1., hold_ratio: float = 0., rng=None):
    """The file selection of BlenderDataset_v2.__init__ (dataset/load_blender.py:270-296): `train_*.npy` are rays of the
    original images, every other `.npy` is pseudo data; pseudo_ratio -1 = everything."""
    rng = rng or np.random
    names = sorted(x for x in os.listdir(datadir) if x.endswith(".npy"))
    pseudo = [f"{datadir}/{x}" for x in names if not x.startswith("train_")]
    original = [f"{datadir}/{x}" for x in names if x.startswith("train_")]
    assert 0 <= pseudo_ratio <= 1 or pseudo_ratio == -1
    if pseudo_ratio == -1 or not original:
        # (with no train_*.npy files the reference's formula draws 0 pseudo shards; its README flow has both kinds)
        all_splits = pseudo + original
    else:
        num_pseudo = int(len(original) / (1. - pseudo_ratio)) - len(original) if pseudo_ratio < 1 else len(pseudo)
        all_splits = rng.choice(pseudo, num_pseudo).tolist() + original
    assert 0 <= hold_ratio < 1
    if hold_ratio > 0:
        all_splits = list(rng.choice(all_splits, int(len(all_splits) * (1 - hold_ratio))))
    return all_splits, len(original), len(pseudo)


class BlenderDataset_v2(torch.utils.data.Dataset):
    """Drop-in for dataset/load_blender.py:257-324: item = (rays_o, rays_d, rgb) of one shard."""

    def __init__(self, datadir, dim_dir=3, dim_rgb=3, rand_crop_size=-1, img_H=0, img_W=0, hold_ratio=0, pseudo_ratio=1.):
        self.datadir = datadir
        self.all_splits, n_orig, n_pseudo = list_shards(datadir, pseudo_ratio, hold_ratio)
        self.dim_dir, self.dim_rgb = dim_dir, dim_rgb
        self.rand_crop_size, self.img_H, self.img_W = rand_crop_size, img_H, img_W
        print(f'Load data done. #All files: {len(self.all_splits)} #Original: {n_orig} #Pseudo: {n_pseudo}')

    def __getitem__(self, index):
        d = torch.from_numpy(np.load(self.all_splits[index]).astype(np.float32, copy=False))
        if self.rand_crop_size > 0:
            x1 = np.random.randint(0, self.img_W - self.rand_crop_size + 1)
            y1 = np.random.randint(0, self.img_H - self.rand_crop_size + 1)
            d = d[y1:y1 + self.rand_crop_size, x1:x1 + self.rand_crop_size, :]
        return d[..., :3], d[..., 3:3 + self.dim_dir], d[..., 3 + self.dim_dir:3 + self.dim_dir + self.dim_rgb]

    def __len__(self):
        return len(self.all_splits)


def read_shards_into(paths, out: torch.Tensor, threads: int = 8) -> None:
    """Payloads of the `.npy` shards `paths` into the rows of `out` (a contiguous float32 HOST tensor, [len(paths) * rows, 9]),
    through the C library's native reader (r2l_read_ray_shards: pread on `threads` threads, the GIL is released)."""
    import ctypes
    from . import _lib
    if out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.numel() % max(len(paths), 1):
        raise ValueError("read_shards_into: `out` must be a contiguous float32 host tensor holding len(paths) equal shards")
    arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    _lib.check(_lib.lib().r2l_read_ray_shards(arr, len(paths), ctypes.c_void_p(out.data_ptr()), out.numel() // max(len(paths), 1),
                                              int(threads)), "r2l_read_ray_shards")


class RayShardLoader:
    """Infinite stream of ray batches: `shards_per_batch` (= --N_rand) shards of `rows` rays each per batch.

    Iterating yields (rays_o, rays_d, target) as [shards_per_batch * rows, 3] HOST views of a pinned ring buffer; the views
    of batch k stay valid until batch k + depth is requested.  `device_batches(device)` yields device tensors instead, with
    the H2D copy of batch k + 1 in flight while the caller works on batch k.  Shard order: a fresh random permutation of this
    rank's shard slice per epoch (InfiniteSamplerWrapper's behaviour, main.py:795-808), from `seed`."""

    def __init__(self, shard_paths, shards_per_batch: int, rows: int = 4096, depth: int = 3, workers: int = 4, seed: int = 0,
                 rank: int = 0, world: int = 1, pin: bool | None = None):
        paths = list(shard_paths)[rank::world]
        if not paths:
            raise ValueError("RayShardLoader: no shards for this rank")
        self.paths, self.shards_per_batch, self.rows = paths, int(shards_per_batch), int(rows)
        self.depth = max(2, int(depth))
        pin = torch.cuda.is_available() if pin is None else pin
        self.buffers = [torch.empty((self.shards_per_batch * self.rows, 9), dtype=torch.float32) for _ in range(self.depth + 1)]
        if pin:
            self.buffers = [b.pin_memory() for b in self.buffers]
        self._rng = np.random.RandomState(seed + 7919 * rank)
        self._order, self._pos = self._rng.permutation(len(paths)), 0
        self._free = queue.Queue()
        for i in range(len(self.buffers)):
            self._free.put(i)
        self._ready = queue.Queue()
        self._workers = max(1, int(workers))
        self._stop = False
        self._held = []   # buffers handed to the consumer, oldest first
        self._thread = threading.Thread(target=self._produce, daemon=True)
        self._thread.start()

    def _next_paths(self):
        out = []
        for _ in range(self.shards_per_batch):
            if self._pos == len(self._order):
                self._order, self._pos = self._rng.permutation(len(self.paths)), 0
            out.append(self.paths[self._order[self._pos]])
            self._pos += 1
        return out

    def _produce(self):
        while not self._stop:
            try:
                i = self._free.get(timeout=0.1)
            except queue.Empty:
                continue
            try:
                read_shards_into(self._next_paths(), self.buffers[i], self._workers)
                self._ready.put((i, None))
            except Exception as e:   # surface I/O errors in the consumer thread
                self._ready.put((i, e))

    def next_buffer(self) -> torch.Tensor:
        """The next [shards_per_batch * rows, 9] pinned batch buffer."""
        while len(self._held) >= self.depth - 1:
            i_old, copied = self._held.pop(0)
            if copied is not None:
                copied.synchronize()      # an async H2D copy still reads the buffer: the producer must not refill it yet
            self._free.put(i_old)
        i, err = self._ready.get()
        if err is not None:
            raise err
        self._held.append([i, None])
        return self.buffers[i]

    def mark_in_flight(self, event) -> None:
        """Tie the buffer returned by the last next_buffer() to a CUDA event recorded after the asynchronous copy that reads
        it: the buffer goes back to the producer only once that event has completed (device_batches does this itself; a
        caller that issues its own non_blocking copies from next_buffer() must, too)."""
        self._held[-1][1] = event

    def __iter__(self):
        return self

    def __next__(self):
        b = self.next_buffer()
        return b[:, :3], b[:, 3:6], b[:, 6:9]

    next = __next__   # the reference calls trainloader.next() (main.py:1305)

    def device_batches(self, device, packed: bool = False):
        """Generator of (rays_o, rays_d, target) DEVICE tensors [N, 3] (column views of one [N, 9] device buffer) - or, with
        packed=True, of the [N, 9] buffer itself (R2LTrainer.step_rays9 / the R2L_INPUT_RAYS9 kernels read the columns in
        place); the copy of the following batch runs on a side stream while the caller consumes the current one."""
        device = torch.device(device)
        copy_stream = torch.cuda.Stream(device)
        dev_bufs = [torch.empty_like(self.buffers[0], device=device) for _ in range(2)]
        events = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def launch(slot):
            host = self.next_buffer()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])        # the step that read this device buffer has finished
                dev_bufs[slot].copy_(host, non_blocking=True)
                events[slot].record(copy_stream)
            self.mark_in_flight(self._copy_done(copy_stream))   # its own event: events[slot] is re-recorded two batches on
        for s in range(2):
            consumed[s].record(torch.cuda.current_stream(device))
        launch(0)
        slot = 0
        while True:
            launch(slot ^ 1)
            torch.cuda.current_stream(device).wait_event(events[slot])
            d = dev_bufs[slot]
            yield d if packed else (d[:, :3], d[:, 3:6], d[:, 6:9])
            consumed[slot].record(torch.cuda.current_stream(device))
            slot ^= 1

    @staticmethod
    def _copy_done(stream):
        ev = torch.cuda.Event()
        ev.record(stream)
        return ev

    def close(self):
        self._stop = True
        self._thread.join(timeout=2)
