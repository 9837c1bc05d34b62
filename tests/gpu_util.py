"""Helpers shared by the -m gpu parity tests (CUDA path vs oracle on identical seeded inputs)."""
import numpy as np
import torch

import oracle
from largesteps_b200 import workloads

DEV = "cuda:0"


def to_dev(v, f, idx_dtype=torch.int64):
    return (torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(DEV),
            torch.from_numpy(np.ascontiguousarray(f)).to(DEV).to(idx_dtype))


def coo_np(M):
    assert M.is_coalesced()
    return M.indices().cpu().numpy(), M.values().cpu().numpy()


def rel_l2(x, y):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))


def fan_mesh(n_rim):
    """One hub vertex joined to n_rim rim vertices: a row with n_rim+1 non-zeros (longer than an SpMM stage)."""
    ang = np.linspace(0, 2 * np.pi, n_rim, endpoint=False)
    v = np.concatenate([[[0, 0, 0.3]], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)]).astype(np.float32)
    i = np.arange(n_rim)
    f = np.stack([np.zeros(n_rim, dtype=np.int64), 1 + i, 1 + (i + 1) % n_rim], 1)
    return v, f


def config1():
    v, f = workloads.icosphere(4)
    return v, f, dict(lambda_=10.0)


def config2(bunny_mesh):
    v, f = bunny_mesh
    v, f = workloads.subdivide(*workloads.subdivide(v, f))
    return v.astype(np.float32), f, dict(lambda_=19.0, cotan=True)


def rhs(r, c, val, V, verts):
    A = oracle.coo_to_scipy(r, c, val, V)
    return workloads.rhs_recipe(lambda x: A @ x, verts)
