"""CPU tests of the host side: ABI surface, module surface, checkpoint format, dispatch guard."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from r2l_b200 import _lib
from r2l_b200 import nerf_raybased as nb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "r2l_b200.h")).read()
    declared = set(re.findall(r"\b(r2l_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.exported_symbols())
    lib = _lib.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.r2l_abi_version() == 1
    assert lib.r2l_packed_bytes() == 1440 * 32768 + (44 + 1 + 43 + 3) * 256 * 4 + 16


def test_c_abi_argument_errors_without_gpu():
    lib = _lib.lib()
    assert lib.r2l_pack_weights(None, None, None) != 0
    assert b"null" in lib.r2l_last_error()
    assert lib.r2l_forward(7, None, None, None, None, None, None, None, None, 0, 5, None) != 0
    assert lib.r2l_forward(0, None, None, None, None, None, None, None, None, 0, 0, None) == 0  # empty batch is a no-op
    assert lib.r2l_backward(0, None, None, None, None, None, None, None, None, 0, -1, None) != 0


def test_state_dict_layout_matches_reference_names(golden_r2l):
    names = [n for n, _, _ in nb.state_dict_layout()]
    assert names == list(golden_r2l["param_names"])
    assert len(names) == 176 and nb.NUM_PARAMS == 5917187


def test_module_state_dict_roundtrip():
    torch.manual_seed(3)
    m = nb.NeRF_v3_2(nb.readme_args(), 1008, 3)
    sd = m.state_dict()
    assert list(sd.keys())[:3] == ["head.0.weight", "head.0.bias", "body.0.body.0.weight"]
    assert sd["head.0.weight"].shape == (256, 1008) and sd["tail.0.bias"].shape == (3,)
    m2 = nb.NeRF_v3_2(nb.readme_args(), 1008, 3)
    m2.load_state_dict(sd)
    assert torch.equal(m.flat, m2.flat)
    # a checkpoint written by the reference (per-layer tensors) loads, too
    ref_like = {k: v.clone() + 1 for k, v in sd.items()}
    m2.load_state_dict(ref_like)
    assert torch.allclose(m2.flat, m.flat + 1)
    with pytest.raises(RuntimeError):
        bad = dict(sd); bad.pop("tail.0.bias"); m2.load_state_dict(bad)


def test_dispatch_guard_names_the_flag():
    for kw, flag in ((dict(netwidth=128), "netwidth"), (dict(netdepth=44), "netdepth"), (dict(linear_tail=True), "linear_tail"),
                     (dict(use_residual=False), "use_residual"), (dict(act="lrelu"), "--act")):
        with pytest.raises(NotImplementedError, match=flag):
            nb.NeRF_v3_2(nb.readme_args(**kw), 1008, 3)
    with pytest.raises(NotImplementedError, match="input_dim"):
        nb.NeRF_v3_2(nb.readme_args(), 6 * 21, 3)


def test_no_cpu_fallback():
    m = nb.NeRF_v3_2(nb.readme_args(), 1008, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(4, 1008))
    from r2l_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.pack_weights(torch.zeros(nb.NUM_PARAMS))


def test_point_sampler_and_embedder_match_golden(golden_r2l):
    g = golden_r2l
    nb.device = torch.device("cpu")
    ps = nb.PointSampler(400, 400, float(g["focal"]), 16, 2.0, 6.0)
    assert np.array_equal(ps.z_vals.numpy(), g["z_vals"])
    assert np.array_equal(ps.dirs[:3, :5].numpy(), g["dirs_corner"])
    ro, rd = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    assert np.array_equal(ps.sample_train(ro, rd, 0).numpy(), g["pts"])
    torch.manual_seed(1)
    assert np.array_equal(ps.sample_train(ro, rd, 1.0).numpy(), g["pts_jit"])
    pts_test = ps.sample_test(torch.from_numpy(g["c2w"])[:3, :4])
    assert np.array_equal(pts_test[:64].numpy(), g["pts_test_first64"])
    lower, diff = ps.jitter_bounds()
    z = lower + diff * torch.from_numpy(g["t_rand"])
    assert np.array_equal((ro[:, None, :] + rd[:, None, :] * z[:, :, None]).reshape(200, -1).numpy(), g["pts_jit"])
    emb = nb.PositionalEmbedder(L=10)
    assert emb.embed_dim == 21
    # sin/cos: ATen's vectorised and scalar paths differ in the last bit depending on how the array is split over threads
    np.testing.assert_allclose(emb(torch.from_numpy(g["pts"])).numpy(), g["x_embed"], rtol=0, atol=5e-7)
