import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "k-diffusion-inverse-problems_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))


@pytest.fixture(scope="session")
def golden_ffhq():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_ffhq.npz"))


@pytest.fixture(scope="session")
def golden_v2():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v2.npz"))


@pytest.fixture(scope="session")
def golden_stsl():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_stsl.npz"))


@pytest.fixture(scope="session")
def golden_samplers():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_samplers.npz"))
