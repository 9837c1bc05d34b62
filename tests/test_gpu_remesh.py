"""GPU: re-parameterisation after a remesh (scripts/main.py:137-169) -- largesteps_b200.remesh.Reparameterizer
rebuilds M, u = M v and the solver for a new connectivity inside one arena; the optimisation loop carries on."""
import gc

import numpy as np
import pytest
import torch

import oracle
from largesteps_b200 import workloads, parameterize
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import from_differential, to_differential
from largesteps_b200.optimize import AdamUniform
from largesteps_b200.remesh import Reparameterizer
from gpu_util import DEV, to_dev, rel_l2

pytestmark = pytest.mark.gpu


def run_steps(M, u, target, steps, lr=0.02):
    opt = AdamUniform([u], lr=lr)
    losses = []
    for _ in range(steps):
        x = from_differential(M, u, "Cholesky")
        loss = ((x - target) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses


def test_swap_connectivity_mid_loop(bunny_mesh):
    v0, f0 = bunny_mesh
    v0 = v0.astype(np.float32)
    lam = 19.0
    rp = Reparameterizer(lambda_=lam)
    tv, tf = to_dev(v0, f0)
    M, u = rp.update(tv, tf)
    # same matrix as compute_matrix, solver primed: from_differential is a pure solve
    Mref = compute_matrix(tv, tf, lam)
    assert torch.equal(M.indices(), Mref.indices()) and torch.equal(M.values(), Mref.values())
    assert (id(M), "Cholesky") in parameterize._cache
    assert rel_l2(from_differential(M, u).cpu().numpy(), v0) < 1e-5
    target = lambda t_: (t_ * (1.0 + 0.15 * torch.sin(9 * t_[:, :1]))).detach()
    u = u.clone().requires_grad_(True)
    l1 = run_steps(M, u, target(tv), 40)
    assert l1[-1] < l1[0]
    # ---- "remesh": new connectivity from the current shape (midpoint subdivision stands in for Botsch-Kobbelt)
    with torch.no_grad():
        v_now = from_differential(M, u).cpu().numpy()
    v1, f1 = workloads.subdivide(v_now, f0)
    v1 = v1.astype(np.float32)
    old_key = (id(M), "Cholesky")
    tv1, tf1 = to_dev(v1, f1)
    M1, u1 = rp.update(tv1, tf1)
    del M
    gc.collect()
    assert old_key not in parameterize._cache and (id(M1), "Cholesky") in parameterize._cache
    r, c, val, V1 = oracle.compute_matrix(v1, f1, lam)
    idx = M1.indices().cpu().numpy()
    assert (idx[0] == r).all() and (idx[1] == c).all() and (M1.values().cpu().numpy() == val).all()
    np.testing.assert_allclose(u1.cpu().numpy(), oracle.to_differential(r, c, val, V1, v1), rtol=0, atol=2e-5 * np.abs(v1).max() * (1 + 6 * lam))
    ds = oracle.DirectSolver(r, c, val, V1)
    b = u1.cpu().numpy()
    assert rel_l2(from_differential(M1, u1).cpu().numpy(), ds.solve(b)) < 1e-5
    u1 = u1.clone().requires_grad_(True)
    l2 = run_steps(M1, u1, target(tv1), 40, lr=0.016)           # step_size *= 0.8 after a remesh (main.py:164)
    assert l2[-1] < l2[0]
    g = torch.randn(V1, 3, device=DEV)
    x = from_differential(M1, u1)
    (x * g).sum().backward()
    assert rel_l2(u1.grad.cpu().numpy(), ds.solve(g.cpu().numpy())) < 1e-5
    # ---- a second remesh of the same size reuses the arena: no new device allocation
    base = rp.arena.buf.data_ptr()
    M2, u2 = rp.update(tv1 * 1.01, tf1)
    assert rp.arena.buf.data_ptr() == base
    assert rel_l2(from_differential(M2, u2).cpu().numpy(), v1 * 1.01) < 1e-5


def test_reparameterizer_options():
    v, f = workloads.plane(120, seed=0)
    tv, tf = to_dev(v, f)
    for kw in (dict(alpha=0.95), dict(lambda_=19.0, cotan=True)):
        rp = Reparameterizer(method="CG", **kw)
        M, u = rp.update(tv, tf)
        assert rel_l2(from_differential(M, u, "CG").cpu().numpy(), v) < 1e-5
        assert isinstance(parameterize._cache[(id(M), "CG")][0], parameterize.ConjugateGradientSolver)
    with pytest.raises(ValueError, match="Unknown solver type"):
        Reparameterizer(method="LU")
