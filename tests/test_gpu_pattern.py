"""GPU parity: the pattern-only matrix copy (4 bytes per entry for matrices whose off-diagonal values are all equal --
the reference's default system matrix I + lambda L with the uniform Laplacian, geometry.py:112-133; csrc/ls_sell_kernel.cuh
"PAT") against the fp64 direct solve, on the shapes that exercise its layout: a wide row, a single CTA, one cluster, the
cooperative grid, irregular valence; a cotangent matrix must stay on the general copy."""
import numpy as np
import pytest
import torch

import oracle
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import to_differential
from largesteps_b200.solvers import PCGSolver
from gpu_util import DEV, to_dev, rel_l2, fan_mesh

pytestmark = pytest.mark.gpu


def cases(bunny_mesh):
    bv, bf = workloads.subdivide(*bunny_mesh)
    return [("fan21", *fan_mesh(21), dict(lambda_=3.0)),                                   # valence 21: the wide-slice loop
            ("plane64", *workloads.plane(64, seed=0), dict(lambda_=19.0)),               # 128 slices: the cooperative grid, one slice per CTA
            ("plane26", *workloads.plane(26, seed=0), dict(lambda_=19.0)),               # 22 slices: single CTA, everything in shared memory
            ("plane40", *workloads.plane(40, seed=0), dict(lambda_=19.0)),               # 50 slices: 50 CTAs of one slice each
            ("plane300", *workloads.plane(300, seed=0), dict(lambda_=19.0)),             # cooperative grid
            ("bunny_x1", bv.astype(np.float32), bf, dict(lambda_=1.0, alpha=0.9)),       # irregular valence
            ("plane300_cot", *workloads.plane(300, seed=0), dict(lambda_=19.0, cotan=True))]   # not uniform: general copy


def test_pattern_only_copy_solves_to_parity(bunny_mesh):
    for name, v, f, kw in cases(bunny_mesh):
        tv, tf = to_dev(v, f)
        M = compute_matrix(tv, tf, **kw)
        s = PCGSolver(M)
        eng = s.describe()["sell_engine"]
        assert eng == (1 if kw.get("cotan") else 2), (name, eng)
        u = to_differential(M, tv)
        x = s.solve(u)
        r, c, val, V = oracle.compute_matrix(np.asarray(v, np.float64), np.asarray(f), **kw)
        xd = oracle.DirectSolver(r, c, val, V).solve(u.cpu().numpy())
        assert rel_l2(x.cpu().numpy(), xd) < 1e-5, name
        x2 = s.solve(u)
        assert torch.equal(x, x2), name                      # deterministic
