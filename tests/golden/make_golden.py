"""Generate tests/golden/*.npz by running the REFERENCE module (imported read-only from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The fixtures pin (a) the oracle restatement in oracle/r2l_oracle.py and (b) the CUDA path, on the same
seeded inputs.  Nothing here is imported by the product.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    sys.path.insert(0, REF)
    saved = sys.modules.pop("model", None), sys.modules.pop("model.nerf_raybased", None)
    import importlib
    mod = importlib.import_module("model.nerf_raybased")
    helpers = importlib.import_module("utils.run_nerf_raybased_helpers")
    torch.autograd.set_detect_anomaly(False)  # the reference switches it on at import (:4)
    sys.path.remove(REF)
    for name in ("model", "model.nerf_raybased", "utils", "utils.run_nerf_raybased_helpers"):
        sys.modules.pop(name, None)
    return mod, helpers


def ref_args():
    trial = types.SimpleNamespace(ON=True, body_arch="resmlp", res_scale=1.0, n_learnable=2, inact="relu",
                                  outact="none", n_block=-1, near=-1, far=-1)
    return types.SimpleNamespace(netdepth=88, netwidth=256, layerwise_netwidths="", act="relu", linear_tail=False,
                                 use_residual=True, trial=trial)


def pose_spherical_np(theta, phi, radius):
    """dataset/load_blender.py:22-28 restated with numpy (camera-to-world of a point on a sphere)."""
    def trans_t(t):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], np.float32)

    def rot_phi(p):
        return np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], np.float32)

    def rot_theta(th):
        return np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], np.float32)

    c2w = trans_t(radius)
    c2w = rot_phi(phi / 180. * np.pi) @ c2w
    c2w = rot_theta(theta / 180. * np.pi) @ c2w
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float32) @ c2w
    return c2w.astype(np.float32)


def flat_from_module(m):
    return torch.cat([v.detach().reshape(-1) for v in m.state_dict().values()]).numpy()


def main():
    ref, helpers = import_reference()
    ref.device = torch.device("cpu")
    torch.set_num_threads(8)
    out = {}

    # ---------------- R2L model, seed 0 ----------------
    torch.manual_seed(0)
    emb = ref.PositionalEmbedder(L=10)
    model = ref.NeRF_v3_2(ref_args(), 16 * 3 * emb.embed_dim, 3)
    names = list(model.state_dict().keys())
    flat = flat_from_module(model)
    assert flat.shape == (5917187,)
    out["param_names"] = np.array(names)
    out["param_sum"] = np.array([float(v.double().sum()) for v in model.state_dict().values()])
    out["param_sumsq"] = np.array([float((v.double() ** 2).sum()) for v in model.state_dict().values()])

    # seeded init in the product must reproduce these weights bit for bit
    from r2l_b200.nerf_raybased import init_flat_params
    mine = init_flat_params(0).numpy()
    assert np.array_equal(mine, flat), "init_flat_params(0) does not reproduce the reference's seed-0 weights"

    # ---------------- rays: lego-like pose, 400x400, focal from main.py:927 ----------------
    H = W = 400
    focal = 555.5555155968841
    near, far = 2.0, 6.0
    sampler = ref.PointSampler(H, W, focal, 16, near, far)
    rng = np.random.RandomState(0)
    c2w = torch.from_numpy(pose_spherical_np(rng.uniform(-180, 180), rng.uniform(-90, 0), 4.0))
    rays_o_full, rays_d_full = helpers.get_rays(H, W, focal, c2w[:3, :4])
    pix = rng.choice(H * W, size=200, replace=False)
    rays_o = rays_o_full.reshape(-1, 3)[pix].contiguous()
    rays_d = rays_d_full.reshape(-1, 3)[pix].contiguous()
    torch.manual_seed(1)
    t_rand = torch.rand(200, 16)
    out.update(c2w=c2w.numpy(), pix=pix, rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), t_rand=t_rand.numpy(),
               z_vals=sampler.z_vals.numpy(), dirs_corner=sampler.dirs[:3, :5].numpy(), focal=np.float64(focal))

    pts = sampler.sample_train(rays_o, rays_d, perturb=0)
    out["pts"] = pts.numpy()
    # sample_train with perturb>0 draws torch.rand internally; replay it with a seeded generator
    torch.manual_seed(1)
    pts_jit = sampler.sample_train(rays_o, rays_d, perturb=1.0)
    out["pts_jit"] = pts_jit.numpy()
    # sample_test on the same pose: first 64 pixels
    pts_test = sampler.sample_test(c2w[:3, :4])
    out["pts_test_first64"] = pts_test[:64].numpy()
    out["pts_test_pix"] = pts_test[pix].numpy()

    x = emb(pts)
    out["x_embed"] = x.numpy()
    with torch.no_grad():
        rgb = model(x)
        rgb_jit = model(emb(pts_jit))
        rgb_test = model(emb(pts_test[pix]))
    out.update(rgb=rgb.numpy(), rgb_jit=rgb_jit.numpy(), rgb_test_pix=rgb_test.numpy())

    # intermediate activations at blocks {0, 9, 19, 29, 42}
    with torch.no_grad():
        h = model.head(x)
        z = h
        for k, blk in enumerate(model.body):
            z = blk(z)
            if k in (0, 9, 19, 29, 42):
                out[f"z_after_block{k}"] = z.numpy()
        out["h_head"] = h.numpy()

    # fp64 evaluation of the same module (truth for tolerances)
    import copy
    m64 = copy.deepcopy(model).double()
    with torch.no_grad():
        out["rgb_f64"] = m64(emb(pts).double()).numpy()

    # ---------------- loss + gradients (main.py:1377, lw_rgb = 1) ----------------
    torch.manual_seed(2)
    target = torch.rand(200, 3)
    out["target"] = target.numpy()
    model.zero_grad()
    loss = ref.img2mse(model(x), target)
    loss.backward()
    g32 = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).numpy()
    m64.zero_grad()
    loss64 = ref.img2mse(m64(x.double()), target.double())
    loss64.backward()
    g64 = torch.cat([p.grad.reshape(-1) for p in m64.parameters()]).numpy()
    idx = np.arange(0, g32.size, 997)
    out.update(loss=np.float32(loss.item()), loss_f64=np.float64(loss64.item()), grad_idx=idx, grad_f32_sub=g32[idx],
               grad_f64_sub=g64[idx], grad_f64_norm=np.float64(np.linalg.norm(g64)),
               grad_f32_vs_f64_rel=np.float64(np.linalg.norm(g32 - g64) / np.linalg.norm(g64)))
    tn64, terr = [], []
    off = 0
    for p in model.parameters():
        n = p.numel()
        tn64.append(np.linalg.norm(g64[off:off + n]))
        terr.append(np.linalg.norm(g32[off:off + n] - g64[off:off + n]) / max(np.linalg.norm(g64[off:off + n]), 1e-300))
        off += n
    out["grad_tensor_norm_f64"] = np.array(tn64)
    out["grad_tensor_ref32_relerr"] = np.array(terr)
    # a few complete small tensors for exact comparisons
    sd_off = {}
    off = 0
    for name, v in model.state_dict().items():
        sd_off[name] = (off, v.numel())
        off += v.numel()
    for name in ("tail.0.weight", "tail.0.bias", "head.0.bias", "body.0.body.0.bias", "body.42.body.2.bias", "body.20.body.0.bias"):
        o, n = sd_off[name]
        out["g64_" + name] = g64[o:o + n]
        out["g32_" + name] = g32[o:o + n]
    np.savez_compressed(os.path.join(HERE, "r2l_seed0.npz"), **out)
    print("r2l_seed0.npz:", {k: getattr(v, "shape", None) for k, v in out.items() if k.startswith(("rgb", "grad_f32_vs"))},
          "ref32 vs f64 flat grad rel:", out["grad_f32_vs_f64_rel"])

    # ---------------- teacher NeRF + raw2outputs + sample_pdf ----------------
    t = {}
    torch.manual_seed(0)
    teacher = ref.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
    t["param_names"] = np.array(list(teacher.state_dict().keys()))
    t["param_sum"] = np.array([float(v.double().sum()) for v in teacher.state_dict().values()])
    embed_fn, ch = ref.get_embedder(10, 0)
    embeddirs_fn, chv = ref.get_embedder(4, 0)
    assert (ch, chv) == (63, 27)
    n_r, n_s = 48, 24
    ro, rd = rays_o[:n_r], rays_d[:n_r]
    viewdirs = rd / torch.norm(rd, dim=-1, keepdim=True)
    z = torch.linspace(0., 1., n_s)
    z = (near * (1. - z) + far * z).expand(n_r, n_s).contiguous()
    pts3 = ro[..., None, :] + rd[..., None, :] * z[..., :, None]
    with torch.no_grad():
        raw = ref.run_network(pts3, viewdirs, teacher, embed_fn, embeddirs_fn, netchunk=1024)
    t.update(rays_o=ro.numpy(), rays_d=rd.numpy(), viewdirs=viewdirs.numpy(), z_vals=z.numpy(), pts=pts3.numpy(), raw=raw.numpy())
    t["embed_pts_first8"] = embed_fn(pts3.reshape(-1, 3)[:8]).numpy()
    # raw2outputs on (a) the network output, (b) a harsher synthetic raw with large densities
    ref_print = lambda *a, **k: None
    for tag, raw_in, wb in (("net", raw, True), ("synth", None, False)):
        if raw_in is None:
            torch.manual_seed(3)
            raw_in = torch.randn(n_r, n_s, 4) * 3.0
            raw_in[0, :, 3] = -1.0   # fully transparent ray: acc = 0, disp = NaN in the reference
        with torch.no_grad():
            rgb_map, disp_map, acc_map, weights, depth_map = ref.raw2outputs(raw_in, z, rd, 0, wb, False, -1, ref_print)
        t[f"r2o_{tag}_raw"] = raw_in.numpy()
        t[f"r2o_{tag}_rgb"] = rgb_map.numpy()
        t[f"r2o_{tag}_disp"] = disp_map.numpy()
        t[f"r2o_{tag}_acc"] = acc_map.numpy()
        t[f"r2o_{tag}_weights"] = weights.numpy()
        t[f"r2o_{tag}_depth"] = depth_map.numpy()
    # sample_pdf (deterministic) on the weights above: utils/create_data.py:505-511
    z_mid = .5 * (z[..., 1:] + z[..., :-1])
    w_in = torch.from_numpy(t["r2o_net_weights"])[..., 1:-1]
    zs = helpers.sample_pdf(z_mid, w_in, 32, det=True)
    t.update(pdf_bins=z_mid.numpy(), pdf_weights=w_in.numpy(), pdf_samples=zs.numpy())
    np.savez_compressed(os.path.join(HERE, "teacher_seed0.npz"), **t)
    print("teacher_seed0.npz written; raw range", float(raw.min()), float(raw.max()))


def make_pose():
    """pose_seed0.npz: the per-frame body of render_path (main.py:300-309,:322-324,:338) on two poses of a small
    non-square frame: PointSampler.sample_test -> PositionalEmbedder -> NeRF_v3_2 (seed 0) -> to8b."""
    ref, _ = import_reference()
    ref.device = torch.device("cpu")
    torch.set_num_threads(8)
    torch.manual_seed(0)
    emb = ref.PositionalEmbedder(L=10)
    model = ref.NeRF_v3_2(ref_args(), 16 * 3 * emb.embed_dim, 3)
    H, W, focal = 18, 20, 27.77777577984421   # the lego focal scaled to a 20-pixel-wide frame
    sampler = ref.PointSampler(H, W, focal, 16, 2.0, 6.0)
    rng = np.random.RandomState(5)
    poses = np.stack([pose_spherical_np(rng.uniform(-180, 180), rng.uniform(-90, 0), 4.0)[:3, :4] for _ in range(2)])
    out = dict(H=np.int64(H), W=np.int64(W), focal=np.float64(focal), c2w=poses, z_vals=sampler.z_vals.numpy(),
               dirs=sampler.dirs.numpy())
    frames, frames8, pts_all = [], [], []
    for c2w in poses:
        pts = sampler.sample_test(torch.from_numpy(c2w))
        with torch.no_grad():
            rgb = model(emb(pts))
        rgb = rgb.view(H, W, 3)
        pts_all.append(pts.numpy())
        frames.append(rgb.numpy())
        frames8.append(ref.to8b(rgb))
    out.update(pts=np.stack(pts_all), rgb=np.stack(frames), rgb8=np.stack(frames8))
    np.savez_compressed(os.path.join(HERE, "pose_seed0.npz"), **out)
    print("pose_seed0.npz written:", out["rgb"].shape, out["rgb8"].dtype, out["rgb8"].reshape(-1, 3)[:3])


def make_camera():
    """camera_seed0.npz: pose_spherical (dataset/load_blender.py:10-28; the module itself needs imageio, so its pose lines are
    executed from source) for a few angles, and get_rays (utils/run_nerf_raybased_helpers.py:231-257) on one of the poses."""
    _, helpers = import_reference()
    src = open(os.path.join(REF, "dataset", "load_blender.py")).read()
    ns = {"torch": torch, "np": np}
    exec(src[src.index("trans_t = lambda"):src.index("def load_blender_data")], ns)
    angles = np.array([[30., -40.], [-170., -5.], [0., -90.], [123.456, -67.89], [-1e-3, -1e-3]])
    c2w = np.stack([ns["pose_spherical"](float(t), float(p), 4.0).numpy() for t, p in angles])
    H, W, focal = 5, 7, 9.5
    helpers.device = torch.device("cpu")
    rays_o, rays_d = helpers.get_rays(H, W, focal, torch.from_numpy(c2w[3][:3, :4]))
    np.savez_compressed(os.path.join(HERE, "camera_seed0.npz"), angles=angles, radius=np.float64(4.0), c2w=c2w, H=np.int64(H),
                        W=np.int64(W), focal=np.float64(focal), rays_o=rays_o.numpy(), rays_d=rays_d.numpy())
    print("camera_seed0.npz written:", c2w.shape, rays_d.shape)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "camera"):
        make_camera()
    if which in ("all", "main"):
        main()
    if which in ("all", "pose"):
        make_pose()
