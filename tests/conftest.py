import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with `pytest -m gpu` on the GPU box")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def dino():
    with np.load(os.path.join(ROOT, "tsbb15-3d-reconstruction-project_b200", "data", "dino_data.npz")) as z:
        d = {k: z[k] for k in z.files}
    d["tracks"] = d.pop("tracks_x100").astype(np.float64) / 100.0
    return d


@pytest.fixture(scope="session")
def f_golden():
    return _load("f_path_golden.npz")


@pytest.fixture(scope="session")
def pnp_golden():
    return _load("pnp_golden.npz")


@pytest.fixture(scope="session")
def gs_golden():
    return _load("gs_golden.npz")


@pytest.fixture(scope="session")
def geom_golden():
    return _load("geom_golden.npz")


@pytest.fixture(scope="session")
def noisy01(dino):
    """Noisy tracked pair (0,1) of imgdata/points.txt as (2,N),(2,N) — the BASELINE config-1 input (N = 257)."""
    tr = dino["tracks"]
    y1, y2 = tr[:, 0:2], tr[:, 2:4]
    ok = np.logical_and(np.any(y1 != -1, axis=1), np.any(y2 != -1, axis=1))
    return np.array(y1[ok]).T.copy(), np.array(y2[ok]).T.copy()


@pytest.fixture(scope="session")
def rg():
    import tsbb15_b200
    return tsbb15_b200
