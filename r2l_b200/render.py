"""Teacher volumetric rendering on the GPU: coarse query -> raw2outputs -> inverse-CDF resampling + sorted merge ->
fine query -> raw2outputs, every stage a kernel of libr2l_b200 and no host round trip in between.

Mirrors `render_rays` / `batchify_rays` / `render` of the reference's pseudo-data generator
(utils/create_data.py:405-544, :80-94, :97-176; the same functions exist in main.py:624-756) — argument names, returned
dictionary keys and numerics — for the path BASELINE.json calls "NeRF teacher pseudo-data generation".  The reference
moves weights to the CPU for sample_pdf (create_data.py:506-511); here it stays on the device."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from . import nerf_raybased as nb


def _pdf_uniforms(shape, det, pytest, device):
    """The uniforms sample_pdf inverts the CDF at (utils/run_nerf_raybased_helpers.py:291-307): None = linspace (det), torch's
    CPU generator otherwise, numpy's generator seeded with 0 under pytest=True (then det means numpy's linspace)."""
    if pytest:
        np.random.seed(0)
        if det:
            return torch.Tensor(np.broadcast_to(np.linspace(0., 1., shape[-1]), shape).copy()).to(device)
        return torch.Tensor(np.random.rand(*shape)).to(device)
    return None if det else torch.rand(shape).to(device)


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """utils/run_nerf_raybased_helpers.py:283-330 on the GPU: bins [N,B], weights [N,B-1] -> samples [N,N_samples]."""
    u = _pdf_uniforms(list(weights.shape[:-1]) + [N_samples], det, pytest, bins.device)
    return ops.sample_pdf(bins, weights, N_samples, u)


def get_rays(H: int, W: int, focal: float, c2w: torch.Tensor):
    """rays_o, rays_d [H,W,3] of utils/run_nerf_raybased_helpers.py:231-257 (trans_origin ''), on c2w's device."""
    dev = c2w.device
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t().to(dev), j.t().to(dev)
    dirs = torch.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs.unsqueeze(dim=-2) * c2w[:3, :3], -1)
    return c2w[:3, -1].expand(rays_d.shape), rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """Rays of a forward-facing scene in normalised device coordinates (utils/run_nerf_raybased_helpers.py:260-280)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def _ray_points(rays_o, rays_d, z_vals):
    """pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None] (utils/create_data.py:486-487, :517) as a lazy
    nb.RayPoints: the fused teacher query builds the points in its prologue, anything else gets the real tensor."""
    if z_vals.is_cuda and z_vals.dtype == torch.float32 and rays_o.dtype == torch.float32:
        return nb.RayPoints(rays_o, rays_d, z_vals)
    return rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False):
    """Volumetric rendering of a ray batch [N, 8 or 11] = (o, d, near, far[, viewdir]); returns the reference's dict.
    pytest=True replaces every random draw by numpy's generator re-seeded with 0 (utils/create_data.py:473-476 and the hooks
    of raw2outputs / sample_pdf), as the reference does."""
    dev = ray_batch.device
    n_rays = ray_batch.shape[0]
    rays_o, rays_d = ray_batch[:, 0:3].contiguous(), ray_batch[:, 3:6].contiguous()
    viewdirs = ray_batch[:, -3:].contiguous() if ray_batch.shape[-1] > 8 else None
    bounds = torch.reshape(ray_batch[..., 6:8], [-1, 1, 2])
    near, far = bounds[..., 0], bounds[..., 1]
    t_vals = torch.linspace(0., 1., steps=N_samples).to(dev)
    if not lindisp:
        z_vals = near * (1. - t_vals) + far * (t_vals)
    else:
        z_vals = 1. / (1. / near * (1. - t_vals) + 1. / far * (t_vals))
    z_vals = z_vals.expand([n_rays, N_samples])
    if perturb > 0.:
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        t_rand = torch.rand(z_vals.shape).to(dev)
        if pytest:
            np.random.seed(0)
            t_rand = torch.Tensor(np.random.rand(*list(z_vals.shape))).to(dev)
        z_vals = lower + (upper - lower) * t_rand
    z_vals = z_vals.contiguous()
    pts = _ray_points(rays_o, rays_d, z_vals)
    raw = network_query_fn(pts, viewdirs, network_fn)
    rgb_map, disp_map, acc_map, weights, depth_map = nb.raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd, pytest=pytest)
    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0 = rgb_map, disp_map, acc_map
        u = _pdf_uniforms([n_rays, N_importance], perturb == 0., pytest, dev)
        z_samples, z_vals = ops.sample_pdf_merge(z_vals, weights, N_importance, u)     # sample_pdf + sort(cat(...))
        pts = _ray_points(rays_o, rays_d, z_vals)
        run_fn = network_fn if network_fine is None else network_fine
        raw = network_query_fn(pts, viewdirs, run_fn)
        rgb_map, disp_map, acc_map, weights, depth_map = nb.raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd, pytest=pytest)
    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map, 'depth_map': depth_map}
    if retraw:
        ret['raw'] = raw
    if N_importance > 0:
        ret['rgb0'], ret['disp0'], ret['acc0'] = rgb_map_0, disp_map_0, acc_map_0
        ret['z_std'] = torch.std(z_samples, dim=-1, unbiased=False)
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """Render rays in chunks (utils/create_data.py:80-94)."""
    parts = {}
    for i in range(0, rays_flat.shape[0], chunk):
        for k, v in render_rays(rays_flat[i:i + chunk], **kwargs).items():
            parts.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in parts.items()}


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """Render a full frame from a pose `c2w` or a ray batch `rays = (rays_o, rays_d)` (utils/create_data.py:97-176)."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, focal, c2w)
    else:
        rays_o, rays_d = rays
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:     # view directions of c2w on the rays of a static camera
            rays_o, rays_d = get_rays(H, W, focal, c2w_staticcam)
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    sh = rays_d.shape
    if ndc:                                # forward-facing scenes
        rays_o, rays_d = ndc_rays(H, W, focal, 1., rays_o, rays_d)
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near, far = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    rays_cat = torch.cat([rays_o, rays_d, near, far] + ([viewdirs] if use_viewdirs else []), -1)
    all_ret = batchify_rays(rays_cat, chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ['rgb_map', 'disp_map', 'acc_map']
    return [all_ret[k] for k in k_extract] + [{k: v for k, v in all_ret.items() if k not in k_extract}]
