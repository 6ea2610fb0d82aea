"""Tensor-level wrappers over the C ABI: torch owns memory and streams, the library does the arithmetic.

Shape / dtype / device / contiguity checks live here (SURVEY.md section 8b "error convention"); the C
side only sees raw device pointers.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

INPUT_RAYS, INPUT_PTS, INPUT_X, INPUT_RAYS9 = 0, 1, 2, 4
NUM_PARAMS = 5917187
N_SAMPLES = 16
IN_DIM = 1008


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda_f32(t: torch.Tensor, name: str, shape_tail=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: r2l_b200 runs on CUDA tensors only (no CPU fallback); got {t.device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if shape_tail is not None and tuple(t.shape[1:]) != tuple(shape_tail):
        raise ValueError(f"{name}: expected shape [N,{','.join(map(str, shape_tail))}], got {tuple(t.shape)}")
    return t.contiguous()


def packed_bytes() -> int:
    return int(_lib.lib().r2l_packed_bytes())


def pack_weights(flat_params: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """flat fp32 parameters (state_dict order) -> tensor-core operand images (uint8 device buffer)."""
    flat_params = _require_cuda_f32(flat_params, "flat_params")
    if flat_params.numel() != NUM_PARAMS:
        raise ValueError(f"flat_params: expected {NUM_PARAMS} floats, got {flat_params.numel()}")
    if out is None:
        out = torch.empty(packed_bytes(), dtype=torch.uint8, device=flat_params.device)
    with torch.cuda.device(flat_params.device):
        _lib.check(_lib.lib().r2l_pack_weights(_ptr(flat_params), _ptr(out), _stream()), "r2l_pack_weights")
    return out


_workspaces: dict = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def forward(packed: torch.Tensor, *, rays_o=None, rays_d=None, z_vals=None, t_rand=None, z_lower=None,
            z_diff=None, pts=None, x=None, rays9=None, out=None) -> torch.Tensor:
    """rgb[N,3] for one of four input forms:
       rays_o,rays_d (+ z_vals host 16-vector; or z_lower,z_diff,t_rand for the stratified jitter),
       rays9[N,9] = (o | d | rgb) shard rows read in place (same z arguments), pts[N,48], or x[N,1008]."""
    L = _lib.lib()
    kind, in0, in1, t_rand, zl, zd = _resolve_inputs(rays_o, rays_d, z_vals, t_rand, z_lower, z_diff, pts, x, rays9)
    n = in0.shape[0]
    dev = in0.device
    if packed.device != dev:
        raise RuntimeError("packed weights live on a different device")
    if out is None:
        out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        wbytes = int(L.r2l_fwd_workspace_bytes(n))
        ws = _workspace(dev, wbytes)
        _lib.check(L.r2l_forward(kind, _ptr(in0), _ptr(in1), _ptr(t_rand), zl, zd, _ptr(packed), _ptr(out),
                                 _ptr(ws), wbytes, n, _stream()), "r2l_forward")
    return out


def render_poses(packed: torch.Tensor, c2w: torch.Tensor, height: int, width: int, focal: float, z_vals,
                 want_rgb: bool = True, want_rgb8: bool = False):
    """Frames for camera poses c2w[P,3,4] (or [3,4]): rays are generated in-kernel (PointSampler.sample_test).
    Returns (rgb[P,H,W,3] float32 or None, rgb8[P,H,W,3] uint8 = to8b(rgb) or None)."""
    c2w = _require_cuda_f32(c2w, "c2w")
    if c2w.dim() == 2:
        c2w = c2w[None]
    if c2w.dim() != 3 or c2w.shape[1] < 3 or c2w.shape[2] != 4:
        raise ValueError(f"c2w: expected [P,3,4] (or [P,4,4]), got {tuple(c2w.shape)}")
    c2w = c2w[:, :3, :].contiguous()
    if not (want_rgb or want_rgb8):
        raise ValueError("render_poses: nothing to compute (want_rgb and want_rgb8 both False)")
    height, width = int(height), int(width)
    if height <= 0 or width <= 0 or not float(focal) > 0:
        raise ValueError("render_poses: height, width and focal must be positive")
    n_poses, dev = c2w.shape[0], c2w.device
    if packed.device != dev:
        raise RuntimeError("packed weights live on a different device")
    rgb = torch.empty((n_poses, height, width, 3), dtype=torch.float32, device=dev) if want_rgb else None
    rgb8 = torch.empty((n_poses, height, width, 3), dtype=torch.uint8, device=dev) if want_rgb8 else None
    zl = (ctypes.c_float * N_SAMPLES)(*[float(v) for v in z_vals])
    L = _lib.lib()
    with torch.cuda.device(dev):
        wbytes = int(L.r2l_fwd_workspace_bytes(n_poses * height * width))
        ws = _workspace(dev, wbytes)
        _lib.check(L.r2l_render_poses(_ptr(c2w), n_poses, height, width, float(focal), zl, _ptr(packed), _ptr(rgb), _ptr(rgb8),
                                      _ptr(ws), wbytes, _stream()), "r2l_render_poses")
    return rgb, rgb8


def selftest_layer(a: torch.Tensor, packed: torch.Tensor, layer: int) -> torch.Tensor:
    a = _require_cuda_f32(a, "a", (256,))
    c = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().r2l_selftest_layer(_ptr(a), _ptr(packed), layer, _ptr(c), _stream()), "r2l_selftest_layer")
    return c


# ------------------------------------------------------------------------------------------------
# training: fused forward that keeps the tensor-core operand images, fused backward
# ------------------------------------------------------------------------------------------------
class TrainContext:
    """Everything r2l_backward needs from the forward pass (device buffers owned by torch)."""
    __slots__ = ("kind", "n", "rgb", "zf", "fwd_saved", "generation", "device_index")


# The saved operand images are large (86 KiB per ray and pass).  By default they live in one grow-only buffer
# per device that the next forward_train() reuses, so a training loop does no cudaMalloc; a context whose
# buffer has been reused refuses to run backward.  forward_train(..., keep=True) gives a private buffer.
_saved_pool: dict = {}
_generation: dict = {}


def _pooled(device, tag: str, nbytes: int) -> torch.Tensor:
    key = (device.index, tag)
    buf = _saved_pool.get(key)
    if buf is None or buf.numel() < nbytes:
        _saved_pool[key] = None
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _saved_pool[key] = buf
    return buf


def _resolve_inputs(rays_o, rays_d, z_vals, t_rand, z_lower, z_diff, pts, x, rays9=None):
    zl = zd = None
    if x is not None:
        kind, in0, in1 = INPUT_X, _require_cuda_f32(x, "x", (IN_DIM,)), None
    elif pts is not None:
        kind, in0, in1 = INPUT_PTS, _require_cuda_f32(pts, "pts", (3 * N_SAMPLES,)), None
    else:
        if rays9 is not None:
            kind, in0, in1 = INPUT_RAYS9, _require_cuda_f32(rays9, "rays9", (9,)), None
        else:
            kind = INPUT_RAYS
            in0 = _require_cuda_f32(rays_o, "rays_o", (3,))
            in1 = _require_cuda_f32(rays_d, "rays_d", (3,))
            if in0.shape != in1.shape:
                raise ValueError("rays_o / rays_d shape mismatch")
        if t_rand is not None:
            t_rand = _require_cuda_f32(t_rand, "t_rand", (N_SAMPLES,))
            if t_rand.shape[0] != in0.shape[0]:
                raise ValueError("t_rand: wrong number of rays")
            zl = (ctypes.c_float * N_SAMPLES)(*[float(v) for v in z_lower])
            zd = (ctypes.c_float * N_SAMPLES)(*[float(v) for v in z_diff])
        else:
            zl = (ctypes.c_float * N_SAMPLES)(*[float(v) for v in z_vals])
    return kind, in0, in1, t_rand, zl, zd


def forward_train(packed: torch.Tensor, *, rays_o=None, rays_d=None, z_vals=None, t_rand=None, z_lower=None,
                  z_diff=None, pts=None, x=None, rays9=None, keep: bool = False, fwd_saved: torch.Tensor | None = None,
                  workspace: torch.Tensor | None = None):
    """Like forward(), but also returns the TrainContext for backward().  fwd_saved / workspace: caller-owned uint8 buffers
    (train_buffer_bytes(n)) instead of the per-device pools - what a captured CUDA graph needs, since pooled buffers may be
    re-allocated by a later, larger call."""
    L = _lib.lib()
    kind, in0, in1, t_rand, zl, zd = _resolve_inputs(rays_o, rays_d, z_vals, t_rand, z_lower, z_diff, pts, x, rays9)
    n, dev = in0.shape[0], in0.device
    ctx = TrainContext()
    ctx.kind, ctx.n = kind, n
    ctx.rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    ctx.zf = torch.empty((n, 256), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nsaved = int(L.r2l_train_fwd_saved_bytes(n))
        if fwd_saved is not None:
            if fwd_saved.numel() < nsaved or fwd_saved.dtype != torch.uint8 or fwd_saved.device != dev:
                raise ValueError(f"fwd_saved: expected a uint8 buffer of >= {nsaved} bytes on {dev}")
            ctx.fwd_saved, ctx.generation = fwd_saved, None
        elif keep:
            ctx.fwd_saved, ctx.generation = torch.empty(nsaved, dtype=torch.uint8, device=dev), None
        else:
            ctx.fwd_saved = _pooled(dev, "fwd", nsaved)
            ctx.generation = _generation[dev.index] = _generation.get(dev.index, 0) + 1
        ctx.device_index = dev.index
        wbytes = int(L.r2l_fwd_workspace_bytes(n))
        ws = _checked_workspace(workspace, dev, wbytes)
        _lib.check(L.r2l_forward_train(kind, _ptr(in0), _ptr(in1), _ptr(t_rand), zl, zd, _ptr(packed), _ptr(ctx.rgb),
                                       _ptr(ctx.zf), _ptr(ctx.fwd_saved), _ptr(ws), wbytes, n, _stream()),
                   "r2l_forward_train")
    return ctx.rgb, ctx


def train_buffer_bytes(n_rays: int):
    """(fwd_saved, bwd_saved, workspace) sizes in bytes for caller-owned buffers of forward_train / backward."""
    L = _lib.lib()
    return int(L.r2l_train_fwd_saved_bytes(n_rays)), int(L.r2l_train_bwd_saved_bytes(n_rays)), int(L.r2l_bwd_workspace_bytes(n_rays))


def _checked_workspace(workspace, dev, nbytes):
    if workspace is None:
        return _workspace(dev, nbytes)
    if workspace.numel() < nbytes or workspace.dtype != torch.uint8 or workspace.device != dev:
        raise ValueError(f"workspace: expected a uint8 buffer of >= {nbytes} bytes on {dev}")
    return workspace


def grad_chunk_ranges(split_layers):
    """[(lo, hi)] float ranges of the flat gradient buffer per chunk of backward(..., split_layers=...), chunk 0 first."""
    n = len(split_layers) + 1
    arr = (ctypes.c_int * max(len(split_layers), 1))(*split_layers)
    out = []
    for c in range(n):
        lo, hi = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.lib().r2l_grad_chunk_range(n, arr, c, ctypes.byref(lo), ctypes.byref(hi)), "r2l_grad_chunk_range")
        out.append((lo.value, hi.value))
    return out


def stream_wait_grad_chunk(chunk: int, stream: "torch.cuda.Stream") -> None:
    """Make `stream` wait until gradient chunk `chunk` of the last chunked backward on this device is complete."""
    _lib.check(_lib.lib().r2l_stream_wait_grad_chunk(int(chunk), ctypes.c_void_p(stream.cuda_stream)), "r2l_stream_wait_grad_chunk")


def backward(packed: torch.Tensor, ctx: TrainContext, grad_rgb: torch.Tensor, grads: torch.Tensor | None = None,
             bwd_saved: torch.Tensor | None = None, workspace: torch.Tensor | None = None, split_layers=None,
             reserve_sms: int = 0) -> torch.Tensor:
    """dL/dparams (flat, state_dict order) for dL/drgb = grad_rgb.  `grads` is overwritten if given.
    split_layers (descending body-layer indices): complete the buffer in len + 1 chunks, top first, with an event behind
    each but the last (stream_wait_grad_chunk) so that a communication stream can reduce them while the backward runs."""
    L = _lib.lib()
    grad_rgb = _require_cuda_f32(grad_rgb, "grad_rgb", (3,))
    if grad_rgb.shape[0] != ctx.n:
        raise ValueError("grad_rgb: wrong number of rays")
    dev = grad_rgb.device
    if ctx.generation is not None and _generation.get(ctx.device_index) != ctx.generation:
        raise RuntimeError("this TrainContext's saved activations were overwritten by a later forward_train(); "
                           "run backward first or use forward_train(..., keep=True)")
    if grads is None:
        grads = torch.empty(NUM_PARAMS, dtype=torch.float32, device=dev)
    elif grads.numel() != NUM_PARAMS or grads.dtype != torch.float32 or not grads.is_contiguous() or grads.device != dev:
        raise ValueError("grads: expected a contiguous float32 CUDA tensor of NUM_PARAMS elements")
    with torch.cuda.device(dev):
        nsaved = int(L.r2l_train_bwd_saved_bytes(ctx.n))
        if bwd_saved is None:
            bwd_saved = _pooled(dev, "bwd", nsaved)
        elif bwd_saved.numel() < nsaved or bwd_saved.dtype != torch.uint8 or bwd_saved.device != dev:
            raise ValueError(f"bwd_saved: expected a uint8 buffer of >= {nsaved} bytes on {dev}")
        wbytes = int(L.r2l_bwd_workspace_bytes(ctx.n))
        ws = _checked_workspace(workspace, dev, wbytes)
        if split_layers:
            arr = (ctypes.c_int * len(split_layers))(*[int(v) for v in split_layers])
            _lib.check(L.r2l_backward_chunked(ctx.kind, _ptr(packed), _ptr(ctx.rgb), _ptr(grad_rgb), _ptr(ctx.zf),
                                              _ptr(ctx.fwd_saved), _ptr(bwd_saved), _ptr(grads), _ptr(ws), wbytes, ctx.n,
                                              _stream(), len(split_layers) + 1, arr, int(reserve_sms)), "r2l_backward_chunked")
        else:
            _lib.check(L.r2l_backward(ctx.kind, _ptr(packed), _ptr(ctx.rgb), _ptr(grad_rgb), _ptr(ctx.zf),
                                      _ptr(ctx.fwd_saved), _ptr(bwd_saved), _ptr(grads), _ptr(ws), wbytes, ctx.n,
                                      _stream()), "r2l_backward")
    return grads


# ------------------------------------------------------------------------------------------------
# compositing and dense encodings
# ------------------------------------------------------------------------------------------------
def raw2outputs(raw: torch.Tensor, z_vals: torch.Tensor, rays_d: torch.Tensor, white_bkgd: bool = False):
    """(rgb_map, disp_map, acc_map, weights, depth_map) of reference raw2outputs with raw_noise_std = 0."""
    raw = _require_cuda_f32(raw, "raw")
    if raw.dim() != 3 or raw.shape[-1] != 4:
        raise ValueError(f"raw: expected [N,S,4], got {tuple(raw.shape)}")
    n, s = raw.shape[0], raw.shape[1]
    z_vals = _require_cuda_f32(z_vals, "z_vals", (s,))
    rays_d = _require_cuda_f32(rays_d, "rays_d", (3,))
    if z_vals.shape[0] != n or rays_d.shape[0] != n:
        raise ValueError("raw / z_vals / rays_d disagree on the number of rays")
    dev = raw.device
    rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    disp, acc, depth = (torch.empty((n,), dtype=torch.float32, device=dev) for _ in range(3))
    weights = torch.empty((n, s), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().r2l_raw2outputs(_ptr(raw), _ptr(z_vals), _ptr(rays_d), n, s, int(bool(white_bkgd)), _ptr(rgb),
                                              _ptr(disp), _ptr(acc), _ptr(weights), _ptr(depth), _stream()), "r2l_raw2outputs")
    return rgb, disp, acc, weights, depth


def positional_embed(x: torch.Tensor, n_freqs: int, style: int) -> torch.Tensor:
    """Dense encoding of x[..., dim]: style 0 = PositionalEmbedder layout, 1 = Embedder (teacher) layout."""
    x = _require_cuda_f32(x, "x")
    dim = x.shape[-1]
    flat = x.reshape(-1, dim).contiguous()
    out = torch.empty((flat.shape[0], dim * (2 * n_freqs + 1)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().r2l_positional_embed(_ptr(flat), _ptr(out), flat.shape[0], dim, n_freqs, style, _stream()),
                   "r2l_positional_embed")
    return out.view(*x.shape[:-1], out.shape[-1])


# ------------------------------------------------------------------------------------------------
# teacher NeRF
# ------------------------------------------------------------------------------------------------
TEACHER_NUM_PARAMS = 595844


def teacher_pack_weights(flat_params: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    flat_params = _require_cuda_f32(flat_params, "flat_params")
    if flat_params.numel() != TEACHER_NUM_PARAMS:
        raise ValueError(f"teacher flat_params: expected {TEACHER_NUM_PARAMS} floats, got {flat_params.numel()}")
    if out is None:
        out = torch.empty(int(_lib.lib().r2l_teacher_packed_bytes()), dtype=torch.uint8, device=flat_params.device)
    with torch.cuda.device(flat_params.device):
        _lib.check(_lib.lib().r2l_teacher_pack_weights(_ptr(flat_params), _ptr(out), _stream()), "r2l_teacher_pack_weights")
    return out


def teacher_forward(packed: torch.Tensor, *, pts=None, viewdirs=None, x_embedded=None) -> torch.Tensor:
    """raw[..., 4] of the teacher MLP.  Either pts[N,S,3] + viewdirs[N,3] (embeddings fused) or x_embedded[P,90]."""
    L = _lib.lib()
    if x_embedded is not None:
        x = _require_cuda_f32(x_embedded, "x_embedded", (90,))
        raw = torch.empty((x.shape[0], 4), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.r2l_teacher_forward(None, None, _ptr(x), _ptr(packed), _ptr(raw), x.shape[0], 1, _stream()), "r2l_teacher_forward")
        return raw
    pts = _require_cuda_f32(pts, "pts")
    if pts.dim() != 3 or pts.shape[-1] != 3:
        raise ValueError(f"pts: expected [N,S,3], got {tuple(pts.shape)}")
    viewdirs = _require_cuda_f32(viewdirs, "viewdirs", (3,))
    if viewdirs.shape[0] != pts.shape[0]:
        raise ValueError("pts / viewdirs disagree on the number of rays")
    n, s = pts.shape[0], pts.shape[1]
    raw = torch.empty((n, s, 4), dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.check(L.r2l_teacher_forward(_ptr(pts), _ptr(viewdirs), None, _ptr(packed), _ptr(raw), n * s, s, _stream()), "r2l_teacher_forward")
    return raw


def teacher_forward_rays(packed: torch.Tensor, rays_o, rays_d, z_vals, viewdirs) -> torch.Tensor:
    """raw[N,S,4] of the teacher MLP at the points rays_o + rays_d * z_vals[N,S], built in the kernel (no pts tensor)."""
    rays_o = _require_cuda_f32(rays_o, "rays_o", (3,))
    rays_d = _require_cuda_f32(rays_d, "rays_d", (3,))
    viewdirs = _require_cuda_f32(viewdirs, "viewdirs", (3,))
    z_vals = _require_cuda_f32(z_vals, "z_vals")
    n = rays_o.shape[0]
    if z_vals.dim() != 2 or z_vals.shape[0] != n or rays_d.shape[0] != n or viewdirs.shape[0] != n:
        raise ValueError("rays_o / rays_d / viewdirs / z_vals disagree on the number of rays")
    s = z_vals.shape[1]
    raw = torch.empty((n, s, 4), dtype=torch.float32, device=rays_o.device)
    with torch.cuda.device(rays_o.device):
        _lib.check(_lib.lib().r2l_teacher_forward_rays(_ptr(rays_o), _ptr(rays_d), _ptr(viewdirs), _ptr(z_vals), _ptr(packed), _ptr(raw),
                                                       n, s, _stream()), "r2l_teacher_forward_rays")
    return raw


def sample_pdf_merge(z_vals: torch.Tensor, weights: torch.Tensor, n_importance: int, u: torch.Tensor | None = None):
    """(z_samples[N,M], z_merged[N,S+M]) : inverse-CDF resampling of the coarse weights + sorted merge, on the GPU.
    u None = deterministic linspace (perturb == 0); else uniforms [N,M] or [M]."""
    z_vals = _require_cuda_f32(z_vals, "z_vals")
    n, s = z_vals.shape
    weights = _require_cuda_f32(weights, "weights", (s,))
    dev = z_vals.device
    if u is None:
        u = torch.linspace(0., 1., steps=n_importance).to(dev)   # evaluated by torch exactly as the reference does
    u = _require_cuda_f32(u, "u")
    if u.dim() == 1 and u.shape[0] == n_importance:
        stride = 0
    elif tuple(u.shape) == (n, n_importance):
        stride = n_importance
    else:
        raise ValueError(f"u: expected [{n_importance}] or [{n},{n_importance}], got {tuple(u.shape)}")
    z_samples = torch.empty((n, n_importance), dtype=torch.float32, device=dev)
    z_merged = torch.empty((n, s + n_importance), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().r2l_sample_pdf_merge(_ptr(z_vals), _ptr(weights), _ptr(u), stride, n, s, n_importance,
                                                   _ptr(z_samples), _ptr(z_merged), _stream()), "r2l_sample_pdf_merge")
    return z_samples, z_merged


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_importance: int, u: torch.Tensor | None = None) -> torch.Tensor:
    """sample_pdf with the reference's arguments: bins[N,B], weights[N,B-1] -> samples[N,n_importance]."""
    bins = _require_cuda_f32(bins, "bins")
    n, b = bins.shape
    weights = _require_cuda_f32(weights, "weights", (b - 1,))
    dev = bins.device
    if u is None:
        u = torch.linspace(0., 1., steps=n_importance).to(dev)
    u = _require_cuda_f32(u, "u")
    if u.dim() == 1 and u.shape[0] == n_importance:
        stride = 0
    elif tuple(u.shape) == (n, n_importance):
        stride = n_importance
    else:
        raise ValueError(f"u: expected [{n_importance}] or [{n},{n_importance}], got {tuple(u.shape)}")
    out = torch.empty((n, n_importance), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().r2l_sample_pdf(_ptr(bins), _ptr(weights), _ptr(u), stride, n, b, n_importance, _ptr(out), _stream()),
                   "r2l_sample_pdf")
    return out


# ------------------------------------------------------------------------------------------------
# loss + gradient in one launch, Adam with device-side step scalars (CUDA-graph friendly train step)
# ------------------------------------------------------------------------------------------------
_loss_scratch: dict = {}


def mse_loss_grad(rgb: torch.Tensor, target: torch.Tensor, grad_scale: float, loss_scale: float, *, grad_rgb=None,
                  per_ray_err=None, loss=None, want_per_ray: bool = False):
    """(loss[1], grad_rgb[N,3], per_ray_err[N] or None): loss = loss_scale * sum((rgb - target)^2), grad_rgb = grad_scale *
    (rgb - target), per_ray_err = mean over the 3 channels of (rgb - target)^2.  img2mse * lw_rgb (main.py:1377) is
    loss_scale = lw_rgb / (3 N), grad_scale = 2 lw_rgb / (3 N_global)."""
    rgb = _require_cuda_f32(rgb, "rgb", (3,))
    if (isinstance(target, torch.Tensor) and target.is_cuda and target.dtype == torch.float32 and target.dim() == 2
            and target.shape[1] == 3 and target.stride(1) == 1 and target.stride(0) >= 3):
        t_stride = target.stride(0)       # e.g. the rgb columns of [N,9] shard rows, read in place
    else:
        target, t_stride = _require_cuda_f32(target, "target", (3,)), 3
    if rgb.shape != target.shape:
        raise ValueError("rgb / target shape mismatch")
    n, dev = rgb.shape[0], rgb.device
    if grad_rgb is None:
        grad_rgb = torch.empty_like(rgb)
    if per_ray_err is None and want_per_ray:
        per_ray_err = torch.empty(n, dtype=torch.float32, device=dev)
    if loss is None:
        loss = torch.empty(1, dtype=torch.float32, device=dev)
    scratch = _loss_scratch.get(dev.index)
    if scratch is None:
        scratch = _loss_scratch[dev.index] = torch.zeros(int(_lib.lib().r2l_loss_scratch_bytes()), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().r2l_mse_loss_grad(_ptr(rgb), _ptr(target), n, int(t_stride), float(grad_scale), float(loss_scale), _ptr(grad_rgb),
                                                _ptr(per_ray_err), _ptr(loss), _ptr(scratch), _stream()), "r2l_mse_loss_grad")
    return loss, grad_rgb, per_ray_err


def adam_hyper(lr: float, beta1: float, beta2: float, step: int, out: torch.Tensor) -> torch.Tensor:
    """The two step-dependent Adam scalars into a HOST float32 tensor of 2 elements (pinned for the graph's H2D copy)."""
    if out.is_cuda or out.dtype != torch.float32 or out.numel() < 2:
        raise ValueError("adam_hyper: `out` must be a host float32 tensor with 2 elements")
    _lib.check(_lib.lib().r2l_adam_hyper(float(lr), float(beta1), float(beta2), int(step), ctypes.c_void_p(out.data_ptr())), "r2l_adam_hyper")
    return out


def adam_schedule_dev(counters: torch.Tensor, hyper: torch.Tensor, lrate: float, lrate_decay: int, warmup_lr, beta1: float, beta2: float):
    """Advance the device-resident counters [global_step, adam_step] (int64 CUDA tensor) and write this iteration's Adam
    scalars into `hyper` (float32 CUDA tensor of >= 3 elements): see r2l_adam_schedule_dev.  Schedule arguments as
    r2l_b200.trainer.lr_at (main.py:1181-1195)."""
    if not (counters.is_cuda and counters.dtype == torch.int64 and counters.numel() >= 2 and counters.is_contiguous()):
        raise RuntimeError("adam_schedule_dev: counters must be a contiguous int64 CUDA tensor of 2 elements")
    if not (hyper.is_cuda and hyper.dtype == torch.float32 and hyper.numel() >= 3 and hyper.is_contiguous()):
        raise RuntimeError("adam_schedule_dev: hyper must be a contiguous float32 CUDA tensor of >= 3 elements")
    start_lr, end_iter = (0.0, 0.0)
    if warmup_lr:
        start_lr, end_iter = [float(x) for x in warmup_lr.split(',')]
    with torch.cuda.device(counters.device):
        _lib.check(_lib.lib().r2l_adam_schedule_dev(float(lrate), start_lr, end_iter, 0.1, float(lrate_decay) * 1000.0, float(beta1),
                                                    float(beta2), _ptr(counters), _ptr(hyper), _stream()), "r2l_adam_schedule_dev")


def adam_step_dev(params, grads, exp_avg, exp_avg_sq, beta1, beta2, eps, hyper_dev):
    for name, t in (("params", params), ("grads", grads), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq), ("hyper", hyper_dev)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError(f"adam_step_dev: {name} must be a contiguous float32 CUDA tensor (no CPU fallback)")
    with torch.cuda.device(params.device):
        _lib.check(_lib.lib().r2l_adam_step_dev(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(), float(beta1),
                                                float(beta2), float(eps), _ptr(hyper_dev), _stream()), "r2l_adam_step_dev")


# ------------------------------------------------------------------------------------------------
# hard-example ray pool (main.py:1325-1347, :1410-1425) as two graph-capturable launches
# ------------------------------------------------------------------------------------------------
def _require_cuda_i32(t, name, numel=None):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise RuntimeError(f"{name}: expected a contiguous int32 CUDA tensor (no CPU fallback)")
    if numel is not None and t.numel() < numel:
        raise ValueError(f"{name}: expected >= {numel} elements, got {t.numel()}")
    return t


def pool_draw(pool_rows: torch.Tensor, pool_state: torch.Tensor, n_out: int, seed: int, counters: torch.Tensor,
              dst_rows: torch.Tensor, slots_out: torch.Tensor) -> None:
    """dst_rows[j] = pool_rows[slots_out[j]] for the first n_out values of the (seed, counters[0])-keyed permutation of the
    pool's slots (r2l_pool_draw).  dst_rows: a contiguous [n_out, 9] view, e.g. the tail rows of the batch buffer."""
    pool_rows = _require_cuda_f32(pool_rows, "pool_rows", (9,))
    if not (dst_rows.is_cuda and dst_rows.dtype == torch.float32 and dst_rows.is_contiguous() and tuple(dst_rows.shape) == (n_out, 9)):
        raise ValueError(f"dst_rows: expected a contiguous float32 CUDA tensor [{n_out}, 9]")
    _require_cuda_i32(pool_state, "pool_state", 1)
    _require_cuda_i32(slots_out, "slots_out", n_out)
    if not (counters.is_cuda and counters.dtype == torch.int64 and counters.is_contiguous()):
        raise RuntimeError("pool_draw: counters must be a contiguous int64 CUDA tensor")
    with torch.cuda.device(pool_rows.device):
        _lib.check(_lib.lib().r2l_pool_draw(_ptr(pool_rows), _ptr(pool_state), int(n_out), int(seed) & (2 ** 64 - 1), _ptr(counters),
                                            _ptr(dst_rows), _ptr(slots_out), _stream()), "r2l_pool_draw")


def pool_update(rays9: torch.Tensor, per_ray_err: torch.Tensor, n_fresh: int, n_hard_in: int, pool_rows: torch.Tensor,
                pool_state: torch.Tensor, slots_out: torch.Tensor | None = None, picked: torch.Tensor | None = None) -> None:
    """The n_hard_in rays of rays9[:n_fresh] with the largest per_ray_err go into the pool: appended at pool_state[0]
    (slots_out None) or over rows slots_out[:n_hard_in] (r2l_pool_update)."""
    rays9 = _require_cuda_f32(rays9, "rays9", (9,))
    per_ray_err = _require_cuda_f32(per_ray_err, "per_ray_err")
    pool_rows = _require_cuda_f32(pool_rows, "pool_rows", (9,))
    if rays9.shape[0] < n_fresh or per_ray_err.numel() < n_fresh:
        raise ValueError("pool_update: rays9 / per_ray_err hold fewer than n_fresh rays")
    _require_cuda_i32(pool_state, "pool_state", 1)
    if slots_out is not None:
        _require_cuda_i32(slots_out, "slots_out", n_hard_in)
    if picked is not None:
        _require_cuda_i32(picked, "picked", n_hard_in)
    with torch.cuda.device(rays9.device):
        _lib.check(_lib.lib().r2l_pool_update(_ptr(rays9), _ptr(per_ray_err), int(n_fresh), int(n_hard_in), _ptr(pool_rows),
                                              _ptr(pool_state), _ptr(slots_out), _ptr(picked), _stream()), "r2l_pool_update")


def pool_slot_host(j: int, size: int, seed: int, step: int) -> int:
    """Slot j of the permutation pool_draw uses (host evaluation of the same function: tests, replaying a run)."""
    v = int(_lib.lib().r2l_pool_slot_host(int(j), int(size), int(seed) & (2 ** 64 - 1), int(step)))
    if v < 0:
        raise ValueError("pool_slot_host: need 0 <= j < size < 2^31")
    return v
