import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def coeff_j():
    from pyemir_b200 import synthetic
    return synthetic.make_config(1)


@pytest.fixture(scope="session")
def scene_j(coeff_j):
    from pyemir_b200 import synthetic
    return synthetic.make_base_scene(coeff_j, seed=1)


@pytest.fixture(scope="session")
def frame_j(coeff_j, scene_j):
    from pyemir_b200 import synthetic
    return synthetic.make_frames(coeff_j, 1, seed=1001, scene=scene_j)[0]


@pytest.fixture()
def rng():
    return np.random.default_rng(12345)
