import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_r2l():
    return dict(np.load(os.path.join(GOLDEN, "r2l_seed0.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_teacher():
    return dict(np.load(os.path.join(GOLDEN, "teacher_seed0.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_pose():
    return dict(np.load(os.path.join(GOLDEN, "pose_seed0.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_surface():
    """tests/golden/make_golden_round2.py: plucker / CNN-style samplers, embed, NCHW forward, load_weights_from_keras."""
    return dict(np.load(os.path.join(GOLDEN, "surface_seed0.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_holes():
    """tests/golden/make_golden_round2.py: raw2outputs with density noise, sample_pdf's pytest hook, ndc_rays - run by the reference."""
    return dict(np.load(os.path.join(GOLDEN, "holes_seed0.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_grad_ref():
    """tests/golden/make_golden_round2.py: the reference's own fp32-vs-fp64 gradient errors at 200 / 1000 / 4096 rays."""
    return dict(np.load(os.path.join(GOLDEN, "grad_ref_seed0.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def keras_weights():
    """The 24 arrays tests/golden/make_golden_round2.py fed to the reference's load_weights_from_keras (same RandomState)."""
    shapes = [(63, 256)] + [(319, 256) if i == 5 else (256, 256) for i in range(1, 8)]
    shapes += [(256, 256), (283, 128), (128, 3), (256, 1)]
    rng = np.random.RandomState(9)
    out = []
    for (i, o) in shapes:
        out += [rng.randn(i, o).astype(np.float32) * 0.05, rng.randn(o).astype(np.float32) * 0.05]
    return out


@pytest.fixture(scope="session")
def flat_seed0():
    """Seed-0 weights; bit-identical to the reference's (asserted when the fixtures were generated and
    re-checked against the stored per-tensor checksums in test_oracle.py)."""
    from r2l_b200.nerf_raybased import init_flat_params
    return init_flat_params(0).numpy()
