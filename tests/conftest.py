import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "large-steps-pytorch_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_assembly():
    return np.load(os.path.join(GOLDEN, "assembly.npz"))


@pytest.fixture(scope="session")
def golden_solve():
    return np.load(os.path.join(GOLDEN, "solve.npz"))


@pytest.fixture(scope="session")
def golden_adam():
    return np.load(os.path.join(GOLDEN, "adam.npz"))


@pytest.fixture(scope="session")
def bunny_mesh():
    d = np.load(os.path.join(GOLDEN, "bunny_mesh.npz"))
    return d["verts"], d["faces"].astype(np.int64)


def rel_l2(x, y):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))
