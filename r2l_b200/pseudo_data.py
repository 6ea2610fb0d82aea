"""Pseudo-data generation with the teacher (BASELINE config 4; utils/create_data.py:780-872, `--create_data rand`):
random camera poses on the upper hemisphere -> rays -> teacher volumetric render (r2l_b200.render, all kernels of
libr2l_b200) -> (o | d | rgb) rows -> shuffled `.npy` shards of 4096 rays, the training input of the R2L network.

Host glue around the hot path: the poses and pixel rays are a few torch ops per frame (bit-compatible with the reference's
pose_spherical / get_rays), everything per sample point runs in the fused kernels, and the rows stay on the device until a
group of frames is written."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import nerf_raybased as nb
from . import render as rr
from .data import write_ray_shards


def pose_spherical(theta: float, phi: float, radius: float) -> torch.Tensor:
    """Camera-to-world [4,4] (float32) of a camera at spherical angles (degrees) looking at the origin
    (dataset/load_blender.py:10-28: trans_t, rot_phi, rot_theta, axis flip; float32 factors, float32 products)."""
    ph, th = phi / 180. * np.pi, theta / 180. * np.pi
    c2w = torch.Tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]]).float()
    c2w = torch.Tensor([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]]).float() @ c2w
    c2w = torch.Tensor([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]]).float() @ c2w
    return torch.Tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w


def get_rand_pose(rng=None, radius: float = 4.0) -> torch.Tensor:
    """A random pose of dataset/load_blender.py:359-368: theta ~ U(-180, 180), phi ~ U(-90, 0), radius 4.  `rng`: a
    numpy RandomState (the reference draws from numpy's global generator)."""
    rng = rng or np.random
    theta = -180 + rng.rand() * 360
    phi = -90 + rng.rand() * 90
    return pose_spherical(theta, phi, radius)


get_rays = rr.get_rays      # utils/run_nerf_raybased_helpers.py:231-257 (lives with render() since round 2)


def render_pseudo_frame(pose, H, W, focal, teacher, teacher_fine, near=2., far=6., N_samples=64, N_importance=128,
                        white_bkgd=True, chunk=1024 * 32, perturb=0.):
    """One frame of pseudo data: rows [H*W, 9] = (o | d | rgb) on the device (utils/create_data.py:815-845).  `teacher`,
    `teacher_fine`: r2l_b200 NeRF modules on CUDA; multires 10 / 4 embedders as create_nerf builds them (:251-293)."""
    embed_fn, _ = nb.get_embedder(10, 0)
    embeddirs_fn, _ = nb.get_embedder(4, 0)

    def query(inputs, viewdirs, network_fn):
        return nb.run_network(inputs, viewdirs, network_fn, embed_fn, embeddirs_fn, netchunk=1024 * 64)
    dev = next(teacher.parameters()).device
    rays_o, rays_d = get_rays(H, W, focal, pose[:3, :4].to(dev))     # (pose[:3,:4], as the reference insists, :820)
    rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
    with torch.no_grad():
        rgb, _, _, _ = rr.render(H, W, focal, chunk=chunk, rays=(rays_o, rays_d), ndc=False, near=near, far=far, use_viewdirs=True,
                                 network_fn=teacher, network_query_fn=query, N_samples=N_samples, N_importance=N_importance,
                                 network_fine=teacher_fine, white_bkgd=white_bkgd, perturb=perturb, raw_noise_std=0.)
    return torch.cat([rays_o, rays_d, rgb], dim=-1).view(H * W, 9)


def generate_pseudo_data(teacher, teacher_fine, datadir: str, n_pose: int, H: int = 400, W: int = 400, focal: float = 555.5555155968841,
                         i_save: int = 100, split_size: int = 4096, seed=None, use_rand_focal: bool = False, **render_kw):
    """The `--create_data rand` loop (utils/create_data.py:807-872): n_pose random poses rendered by the teacher, every i_save
    frames the accumulated rows are shuffled twice and written as `data_{k}.npy` shards of split_size rays (the tail that does
    not fill a shard is dropped, as in the reference).  Existing shards in `datadir` are kept and numbering continues (:790-797).
    Returns the number of shards written."""
    os.makedirs(datadir, exist_ok=True)
    rng = np.random.RandomState(seed) if seed is not None else np.random
    split = len([x for x in os.listdir(datadir) if x.endswith(".npy")])
    written, rows = 0, []
    for i in range(1, n_pose + 1):
        pose = get_rand_pose(rng)
        focal_ = focal * (rng.rand() + 1) if use_rand_focal else focal          # scale focal by [1, 2) (:816-818)
        rows.append(render_pseudo_frame(pose, H, W, focal_, teacher, teacher_fine, **render_kw))
        if i % i_save == 0 or i == n_pose:
            data = torch.cat(rows, dim=0).cpu().numpy()
            paths = write_ray_shards(data, datadir, split_size, first_index=split + 1, rng=rng)
            split += len(paths)
            written += len(paths)
            rows = []
    return written
