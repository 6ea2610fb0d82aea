"""CPU restatement (numpy) of the R2L hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module; the product (r2l_b200/) never does and fails loudly when its CUDA library is missing.

Every function cites the reference lines it restates (paths relative to /root/reference).  The oracle is
pinned against the reference itself: tests/golden/make_golden.py imports the reference module in the build
container, runs it on seeded inputs and stores inputs/outputs in tests/golden/*.npz; tests/test_oracle.py
checks this file against those fixtures (the reference ships no tests or golden vectors of its own,
SURVEY.md section 4 / 8c).

All functions are dtype-generic: pass float32 arrays for the reference's arithmetic, float64 for "truth".
"""
from __future__ import annotations

import numpy as np

# ---- geometry of the README configuration (README.md:51; option.py:93-97) ----
WIDTH = 256
N_BLOCKS = 43
N_SAMPLES = 16
N_FREQS = 10
EMBED = 2 * N_FREQS + 1
IN_DIM = N_SAMPLES * 3 * EMBED  # 1008
OFF_HEAD_W = 0
OFF_HEAD_B = WIDTH * IN_DIM
OFF_BODY = OFF_HEAD_B + WIDTH
LINEAR_STRIDE = WIDTH * WIDTH + WIDTH
OFF_TAIL_W = OFF_BODY + 2 * N_BLOCKS * LINEAR_STRIDE
OFF_TAIL_B = OFF_TAIL_W + 3 * WIDTH
NUM_PARAMS = OFF_TAIL_B + 3  # 5,917,187


def unflatten_params(flat: np.ndarray):
    """Flat state_dict-ordered buffer -> dict of views (NeRF_v3_2.__init__, model/nerf_raybased.py:483-537)."""
    assert flat.shape == (NUM_PARAMS,)
    p = {
        "head_w": flat[OFF_HEAD_W:OFF_HEAD_B].reshape(WIDTH, IN_DIM),
        "head_b": flat[OFF_HEAD_B:OFF_BODY],
        "body": [],
        "tail_w": flat[OFF_TAIL_W:OFF_TAIL_B].reshape(3, WIDTH),
        "tail_b": flat[OFF_TAIL_B:NUM_PARAMS],
    }
    for l in range(2 * N_BLOCKS):
        o = OFF_BODY + l * LINEAR_STRIDE
        p["body"].append((flat[o:o + WIDTH * WIDTH].reshape(WIDTH, WIDTH), flat[o + WIDTH * WIDTH:o + LINEAR_STRIDE]))
    return p


# ------------------------------------------------------------------------------------------------
# PointSampler  (model/nerf_raybased.py:76-126)
# ------------------------------------------------------------------------------------------------
def torch_linspace(start, end, steps: int, dtype=np.float32) -> np.ndarray:
    """torch.linspace as ATen evaluates it on CPU: step = (end-start)/(steps-1) in `dtype`; the first half
    is start + step*i, the second half end - step*(steps-1-i), each with ONE rounding (vectorised FMA).
    The reference calls torch.linspace at model/nerf_raybased.py:79-81,:88 — torch is the pinned dependency
    whose arithmetic this restates (checked bit-exactly against torch in tests/test_oracle.py)."""
    if dtype != np.float32:
        return np.linspace(start, end, steps).astype(dtype)
    s, e = np.float32(start), np.float32(end)
    step = np.float64(np.float32((e - s) / np.float32(steps - 1)))
    i = np.arange(steps, dtype=np.float64)
    lo = (np.float64(s) + step * i).astype(np.float32)
    hi = (np.float64(e) - step * (steps - 1 - i)).astype(np.float32)
    return np.where(np.arange(steps) < steps // 2, lo, hi)


def sampler_dirs(H: int, W: int, focal: float, dtype=np.float32) -> np.ndarray:
    """:80-86 — pixel directions [H,W,3] = [(i-W/2)/f, -(j-H/2)/f, -1]."""
    i, j = np.meshgrid(torch_linspace(0, W - 1, W, dtype), torch_linspace(0, H - 1, H, dtype), indexing="xy")
    f = dtype(focal)
    return np.stack([(i - dtype(W * .5)) / f, -(j - dtype(H * .5)) / f, -np.ones_like(i)], axis=-1).astype(dtype)


def sampler_z_vals(near: float, far: float, n_sample: int = N_SAMPLES, dtype=np.float32) -> np.ndarray:
    """:88-90 — z = near*(1-t) + far*t with t = linspace(0,1,n) evaluated in `dtype`."""
    t = torch_linspace(0., 1., n_sample, dtype)
    return (dtype(near) * (dtype(1) - t) + dtype(far) * t).astype(dtype)


def sample_test(dirs: np.ndarray, c2w: np.ndarray, z_vals: np.ndarray) -> np.ndarray:
    """:94-102 — rays from a pose, then pts = o + d*z flattened to [H*W, n_sample*3]."""
    rays_d = np.sum(dirs[..., None, :] * c2w[:3, :3], axis=-1).reshape(-1, 3)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[None, :, None]
    return pts.reshape(pts.shape[0], -1)


def to8b(x: np.ndarray) -> np.ndarray:
    """:16 — (255 * clip(x, 0, 1)).astype(uint8): the product is rounded in x's dtype, the cast truncates."""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def render_poses(flat: np.ndarray, c2w: np.ndarray, H: int, W: int, focal: float, near: float, far: float):
    """The per-frame body of render_path for the R2L network (main.py:300-309 sample_test -> positional_embedder ->
    model; :322-324 view as [H,W,3]; :338 to8b) for poses c2w[P,3,4] -> (rgb[P,H,W,3], rgb8[P,H,W,3])."""
    dt = flat.dtype.type
    dirs, z = sampler_dirs(H, W, focal, dt), sampler_z_vals(near, far, N_SAMPLES, dt)
    frames = [r2l_forward(flat, positional_embed(sample_test(dirs, m.astype(dt), z))).reshape(H, W, 3) for m in c2w]
    rgb = np.stack(frames)
    return rgb, to8b(rgb)


def sample_train(rays_o: np.ndarray, rays_d: np.ndarray, z_vals: np.ndarray, t_rand: np.ndarray | None) -> np.ndarray:
    """:114-126 — t_rand None == perturb 0; otherwise the stratified jitter with the supplied uniforms."""
    z = np.broadcast_to(z_vals[None, :], (rays_o.shape[0], z_vals.shape[0]))
    if t_rand is not None:
        mids = z.dtype.type(.5) * (z[..., 1:] + z[..., :-1])
        upper = np.concatenate([mids, z[..., -1:]], -1)
        lower = np.concatenate([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    return pts.reshape(pts.shape[0], -1)


def jitter_bounds(z_vals: np.ndarray):
    """`lower` and `upper - lower` of :118-121, the two 16-vectors the C ABI takes."""
    mids = z_vals.dtype.type(.5) * (z_vals[1:] + z_vals[:-1])
    upper = np.concatenate([mids, z_vals[-1:]])
    lower = np.concatenate([z_vals[:1], mids])
    return lower, upper - lower


# ------------------------------------------------------------------------------------------------
# PositionalEmbedder  (model/nerf_raybased.py:191-208)
# ------------------------------------------------------------------------------------------------
def positional_embed(x: np.ndarray, L: int = N_FREQS) -> np.ndarray:
    """:198-208 — per coordinate [sin(x 2^0..2^(L-1)), cos(...), x] -> [N, dim*(2L+1)]."""
    w = (2.0 ** np.linspace(0, L - 1, L)).astype(x.dtype)
    y = x[..., None] * w
    y = np.concatenate([np.sin(y), np.cos(y), x[..., None]], axis=-1)
    return y.reshape(y.shape[0], -1)


# ------------------------------------------------------------------------------------------------
# NeRF_v3_2 forward / backward  (model/nerf_raybased.py:443-465, :539-544; loss main.py:1377)
# ------------------------------------------------------------------------------------------------
def sigmoid(x):
    return 1. / (1. + np.exp(-x))


def r2l_forward(flat: np.ndarray, x: np.ndarray, keep: bool = False):
    """x[N,1008] -> rgb[N,3].  head Linear+ReLU; 43x (z + W2 relu(W1 z + b1) + b2); +h; tail Linear+Sigmoid."""
    p = unflatten_params(flat)
    h = np.maximum(x @ p["head_w"].T + p["head_b"], 0)           # :542
    z = h
    saved = {"x": x, "h": h, "z": [], "a": []}
    for k in range(N_BLOCKS):
        (w1, b1), (w2, b2) = p["body"][2 * k], p["body"][2 * k + 1]
        a = np.maximum(z @ w1.T + b1, 0)                           # ResMLP.body :453-456
        if keep:
            saved["z"].append(z)
            saved["a"].append(a)
        z = (a @ w2.T + b2) + z                                    # :462 (res_scale = 1)
    zf = z + h                                                     # :543 use_residual
    rgb = sigmoid(zf @ p["tail_w"].T + p["tail_b"])               # :544 tail Linear + Sigmoid
    if keep:
        saved["zf"] = zf
        saved["rgb"] = rgb
        return rgb, saved
    return rgb


def r2l_loss_and_grads(flat: np.ndarray, x: np.ndarray, target: np.ndarray, lw_rgb: float = 1.0):
    """loss = mean((rgb-target)^2)*lw_rgb (img2mse, nerf_raybased.py:18; main.py:1377) and d loss / d flat."""
    p = unflatten_params(flat)
    rgb, s = r2l_forward(flat, x, keep=True)
    n = x.shape[0]
    dt = flat.dtype.type
    loss = np.mean((rgb - target) ** 2) * dt(lw_rgb)
    g = np.zeros_like(flat)
    gp = unflatten_params(g)
    d_rgb = dt(2.0 * lw_rgb / (3 * n)) * (rgb - target)
    d_logit = d_rgb * rgb * (1 - rgb)
    gp["tail_w"][...] = d_logit.T @ s["zf"]
    gp["tail_b"][...] = d_logit.sum(0)
    g_zf = d_logit @ p["tail_w"]
    gz = g_zf.copy()
    for k in reversed(range(N_BLOCKS)):
        (w1, _), (w2, _) = p["body"][2 * k], p["body"][2 * k + 1]
        (gw1, gb1), (gw2, gb2) = gp["body"][2 * k], gp["body"][2 * k + 1]
        gw2[...] = gz.T @ s["a"][k]
        gb2[...] = gz.sum(0)
        dh = (gz @ w2) * (s["a"][k] > 0)
        gw1[...] = dh.T @ s["z"][k]
        gb1[...] = dh.sum(0)
        gz = gz + dh @ w1
    gh = (gz + g_zf) * (s["h"] > 0)
    gp["head_w"][...] = gh.T @ x
    gp["head_b"][...] = gh.sum(0)
    per_ray_sqerr = ((rgb - target) ** 2).mean(-1)
    return loss, g, rgb, per_ray_sqerr


# ------------------------------------------------------------------------------------------------
# Teacher: Embedder / NeRF / run_network / raw2outputs  (model/nerf_raybased.py:23-73, :226-401)
# ------------------------------------------------------------------------------------------------
def teacher_embed(x: np.ndarray, multires: int) -> np.ndarray:
    """Embedder.embed :54-55 with get_embedder kwargs :62-69: [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]."""
    outs = [x]
    for f in (2.0 ** np.linspace(0., multires - 1, multires)).astype(x.dtype):
        outs += [np.sin(x * f), np.cos(x * f)]
    return np.concatenate(outs, -1)


def teacher_param_shapes(D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,)):
    """state_dict order of NeRF(use_viewdirs=True) :357-375: pts_linears.i, views_linears.0, feature_linear,
    alpha_linear, rgb_linear (weight then bias each)."""
    shapes = []
    for i in range(D):
        fan_in = input_ch if i == 0 else (W + input_ch if (i - 1) in skips else W)
        shapes += [(W, fan_in), (W,)]
    shapes += [(W // 2, input_ch_views + W), (W // 2,)]
    shapes += [(W, W), (W,), (1, W), (1,), (3, W // 2), (3,)]
    return shapes


def teacher_forward(params: list, x: np.ndarray, D=8, input_ch=63, skips=(4,)) -> np.ndarray:
    """NeRF.forward :377-401 (use_viewdirs=True).  params = list of arrays in teacher_param_shapes order."""
    pts, views = x[..., :input_ch], x[..., input_ch:]
    h = pts
    for i in range(D):
        h = np.maximum(h @ params[2 * i].T + params[2 * i + 1], 0)
        if i in skips:
            h = np.concatenate([pts, h], -1)
    o = 2 * D
    vw, vb, fw, fb, aw, ab, rw, rb = params[o:o + 8]
    alpha = h @ aw.T + ab
    feat = h @ fw.T + fb
    hv = np.maximum(np.concatenate([feat, views], -1) @ vw.T + vb, 0)
    rgb = hv @ rw.T + rb
    return np.concatenate([rgb, alpha], -1)


def run_network(inputs: np.ndarray, viewdirs: np.ndarray, params: list, multires=10, multires_views=4) -> np.ndarray:
    """run_network :312-334: embed points [N,S,3] and broadcast view dirs [N,3], apply the MLP, reshape [N,S,4]."""
    flat = inputs.reshape(-1, inputs.shape[-1])
    emb = teacher_embed(flat, multires)
    dirs = np.broadcast_to(viewdirs[:, None], inputs.shape).reshape(-1, 3)
    emb = np.concatenate([emb, teacher_embed(dirs, multires_views)], -1)
    out = teacher_forward(params, emb)
    return out.reshape(list(inputs.shape[:-1]) + [out.shape[-1]])


def raw_noise_pytest(shape, raw_noise_std: float) -> np.ndarray:
    """The deterministic density noise of raw2outputs(pytest=True) (:267-270): np.random.seed(0); rand(*shape) * std, as fp32."""
    np.random.seed(0)
    return (np.random.rand(*shape) * raw_noise_std).astype(np.float32)


def raw2outputs(raw: np.ndarray, z_vals: np.ndarray, rays_d: np.ndarray, white_bkgd: bool = False, noise: np.ndarray | None = None):
    """raw2outputs :226-295; `noise` [N,S] = the density noise of raw_noise_std > 0 (:262-272, added to raw[..., 3] before the
    relu), None for raw_noise_std = 0.  Returns rgb_map, disp_map, acc_map, weights, depth_map."""
    dt = raw.dtype.type
    if noise is not None:
        raw = raw.copy()
        raw[..., 3] = raw[..., 3] + noise.astype(raw.dtype)
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = np.concatenate([dists, np.broadcast_to(dt(1e10), dists[..., :1].shape)], -1)       # :249-252
    dists = dists * np.linalg.norm(rays_d[..., None, :], axis=-1).astype(raw.dtype)            # :255-257
    rgb = sigmoid(raw[..., :3])                                                               # :259
    with np.errstate(over="ignore"):
        alpha = dt(1.) - np.exp(-np.maximum(raw[..., 3], 0) * dists)                          # :246,:272
    trans = np.cumprod(np.concatenate([np.ones((alpha.shape[0], 1), raw.dtype), dt(1.) - alpha + dt(1e-10)], -1), -1)[:, :-1]
    weights = alpha * trans                                                                   # :281-284
    rgb_map = np.sum(weights[..., None] * rgb, -2)
    depth_map = np.sum(weights * z_vals, -1)
    acc_map = np.sum(weights, -1)
    with np.errstate(invalid="ignore", divide="ignore"):
        disp_map = dt(1.) / np.maximum(dt(1e-10), depth_map / acc_map)                        # :288-289 (torch.max: NaN propagates)
        nan = np.isnan(depth_map / acc_map)
    disp_map = np.where(nan, dt(np.nan), disp_map)
    if white_bkgd:
        rgb_map = rgb_map + (dt(1.) - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map


def sample_pdf(bins: np.ndarray, weights: np.ndarray, n_samples: int, u: np.ndarray | None = None) -> np.ndarray:
    """sample_pdf utils/run_nerf_raybased_helpers.py:283-330; u None == det=True (linspace)."""
    dt = weights.dtype.type
    weights = weights + dt(1e-5)
    pdf = weights / np.sum(weights, -1, keepdims=True)
    cdf = np.cumsum(pdf, -1)
    cdf = np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = np.broadcast_to(np.linspace(0., 1., n_samples).astype(weights.dtype), list(cdf.shape[:-1]) + [n_samples])
    inds = np.stack([np.searchsorted(cdf[i], u[i], side="right") for i in range(cdf.shape[0])])
    below = np.maximum(0, inds - 1)
    above = np.minimum(cdf.shape[-1] - 1, inds)
    cdf_g0 = np.take_along_axis(cdf, below, -1)
    cdf_g1 = np.take_along_axis(cdf, above, -1)
    bins_g0 = np.take_along_axis(bins, below, -1)
    bins_g1 = np.take_along_axis(bins, above, -1)
    denom = cdf_g1 - cdf_g0
    denom = np.where(denom < dt(1e-5), np.ones_like(denom), denom)
    t = (u - cdf_g0) / denom
    return bins_g0 + t * (bins_g1 - bins_g0)


def render_rays(rays_o, rays_d, viewdirs, near, far, params_coarse, params_fine, n_samples=64, n_importance=128,
                white_bkgd=True):
    """render_rays utils/create_data.py:405-544 with perturb = 0, lindisp False, raw_noise_std 0: coarse pass, deterministic
    inverse-CDF resampling (:503-511), sorted merge (:513-515), fine pass."""
    dt = rays_o.dtype.type
    t = torch_linspace(0., 1., n_samples, rays_o.dtype.type)
    z = np.broadcast_to(dt(near) * (dt(1.) - t) + dt(far) * t, (rays_o.shape[0], n_samples)).astype(rays_o.dtype)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    raw = run_network(pts, viewdirs, params_coarse)
    rgb0, disp0, acc0, weights, depth0 = raw2outputs(raw, z, rays_d, white_bkgd)
    z_mid = dt(.5) * (z[..., 1:] + z[..., :-1])
    u = np.broadcast_to(torch_linspace(0., 1., n_importance, rays_o.dtype.type), (rays_o.shape[0], n_importance))
    z_samples = sample_pdf(z_mid, weights[..., 1:-1], n_importance, u)
    z_all = np.sort(np.concatenate([z, z_samples], -1), -1)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_all[..., :, None]
    raw = run_network(pts, viewdirs, params_fine)
    rgb, disp, acc, w, depth = raw2outputs(raw, z_all, rays_d, white_bkgd)
    return {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth, "rgb0": rgb0, "disp0": disp0,
            "acc0": acc0, "z_samples": z_samples, "z_vals": z_all, "weights0": weights}


# ------------------------------------------------------------------------------------------------
# training-loop pieces around the network (main.py:1176-1425)
# ------------------------------------------------------------------------------------------------
def lr_schedule(global_step: int, lrate: float, lrate_decay: int, warmup_lr: str | None = None) -> float:
    """main.py:1181-1195 — exponential decay by 0.1 every lrate_decay*1000 steps, optional linear warm-up 'start_lr,end_iter'."""
    decay_rate = 0.1
    decay_steps = lrate_decay * 1000
    if warmup_lr:
        start_lr, end_iter = [float(x) for x in warmup_lr.split(',')]
        if global_step < end_iter:
            return (lrate - start_lr) / end_iter * global_step + start_lr
        return lrate * (decay_rate ** ((global_step - end_iter) / decay_steps))
    return lrate * (decay_rate ** (global_step / decay_steps))


def img2mse(x: np.ndarray, y: np.ndarray):
    """model/nerf_raybased.py:18."""
    return np.mean((x - y) ** 2)


def per_ray_error(rgb: np.ndarray, target: np.ndarray) -> np.ndarray:
    """main.py:1411-1413 — torch.mean((rgb - target_s)**2, dim=1)."""
    return np.mean((rgb - target) ** 2, axis=1)


def hard_pool_update(hard_rays, hard_pool_full, rays_o, rays_d, target, rgb, batch_size, n_hard_in, hard_mul, rand_ix_out=None):
    """main.py:1410-1425 — the n_hard_in rays of the fresh batch with the largest error replace the first n_hard_in drawn slots
    (pool full) or are appended; the pool is full once it holds batch_size*hard_mul rays.  Returns (hard_rays, full)."""
    err = per_ray_error(rgb[:batch_size], target[:batch_size])
    indices = np.argsort(err, kind="stable")
    hard_indices = indices[-n_hard_in:]
    hard_rays_ = np.concatenate([rays_o[hard_indices], rays_d[hard_indices], target[hard_indices]], axis=-1)
    if hard_pool_full:
        hard_rays = hard_rays.copy()
        hard_rays[rand_ix_out[:n_hard_in]] = hard_rays_
    else:
        hard_rays = np.concatenate([hard_rays, hard_rays_], axis=0)
        if hard_rays.shape[0] >= batch_size * hard_mul:
            hard_pool_full = True
    return hard_rays, hard_pool_full


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (the reference's optimizer, main.py:465,:1406) for one tensor, defaults otherwise; step counts from 1."""
    m = m + (g - m) * (1 - beta1)
    v = v * beta2 + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    p = p - (lr / bc1) * (m / (np.sqrt(v) / np.sqrt(bc2) + eps))
    return p, m, v


# ------------------------------------------------------------------------------------------------
# camera poses and rays of the pseudo-data generator (dataset/load_blender.py:10-28,:359-368; get_rays)
# ------------------------------------------------------------------------------------------------
def pose_spherical(theta: float, phi: float, radius: float) -> np.ndarray:
    """dataset/load_blender.py:10-28 — camera-to-world [4,4] of a camera on a sphere looking at the origin; every factor is a
    float32 matrix (torch.Tensor(...)) and the products are float32 matmuls, as in the reference."""
    f = np.float32
    t = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], f)
    ph, th = phi / 180. * np.pi, theta / 180. * np.pi
    rp = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]], f)
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], f)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], f)
    return (flip @ (rt @ (rp @ t))).astype(f)


def get_rays(H: int, W: int, focal: float, c2w: np.ndarray):
    """utils/run_nerf_raybased_helpers.py:231-257 (no origin translation) — rays_o, rays_d [H,W,3]: the same directions and
    rotation as PointSampler (model/nerf_raybased.py:80-86,:95-99)."""
    dirs = sampler_dirs(H, W, focal, c2w.dtype.type)
    rays_d = np.sum(dirs[..., None, :] * c2w[:3, :3], axis=-1)
    return np.broadcast_to(c2w[:3, -1], rays_d.shape).copy(), rays_d
