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


@pytest.fixture(scope="session")
def golden_traj_small():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_traj_small.npz"))


@pytest.fixture(scope="session")
def golden_full():
    """256x256 goldens of every BASELINE config (tests/golden/make_golden_full.py): one dict over the three files."""
    out = {}
    for part in ("ffhq", "imagenet", "traj"):
        with np.load(os.path.join(ROOT, "tests", "golden", f"golden_full_{part}.npz")) as z:
            out.update({k: z[k] for k in z.files})
    return out


# ---- parity record: every GPU parity test reports its measured errors; the session writes them to gpurun_out/parity_r2.json
# (copied to profiles/parity_r2.json and committed: what e_max / e_l2 actually were on the B200, not just pass / fail) ----------
_PARITY = {}


@pytest.fixture(scope="session")
def parity_log():
    def rec(name, **metrics):
        _PARITY[name] = {k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in metrics.items()}
    return rec


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        path = os.path.join(out_dir, "parity_r2.json")
        old = {}
        if os.path.exists(path):
            try:
                old = json.load(open(path))
            except Exception:
                old = {}
        old.update(_PARITY)
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass
