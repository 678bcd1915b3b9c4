import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden1d():
    return np.load(os.path.join(GOLDEN, "ref1d.npz"))


@pytest.fixture(scope="session")
def golden2d():
    return np.load(os.path.join(GOLDEN, "ref2d.npz"))


@pytest.fixture(scope="session")
def golden2d_post():
    return np.load(os.path.join(GOLDEN, "ref2d_post.npz"))


def golden_cases(g):
    """[(idx, N, k, L)] with k restored to int where the generator used an int."""
    out = []
    for idx, (N, k, L) in enumerate(g["cases"]):
        k = int(k) if float(k).is_integer() and idx in (0, 2) else float(k)
        out.append((idx, int(N), k, float(L)))
    return out


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))
