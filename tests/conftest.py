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
def flat_seed0():
    """Seed-0 weights; bit-identical to the reference's (asserted when the fixtures were generated and
    re-checked against the stored per-tensor checksums in test_oracle.py)."""
    from r2l_b200.nerf_raybased import init_flat_params
    return init_flat_params(0).numpy()
