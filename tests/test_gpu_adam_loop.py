"""GPU parity: fused AdamUniform (csrc/ls_adam.cu) vs the reference's outputs, and the Tutorial-shaped
optimisation loop (config 5 stand-in: from_differential -> loss -> backward -> AdamUniform) vs a CPU oracle loop."""
import numpy as np
import pytest
import torch

import oracle
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import to_differential, from_differential
from largesteps_b200.optimize import AdamUniform
from gpu_util import DEV, to_dev, rel_l2

pytestmark = pytest.mark.gpu


def test_adam_uniform_vs_reference_golden(golden_adam):
    a = golden_adam
    p = torch.nn.Parameter(torch.from_numpy(a["p0"].copy()).to(DEV))
    opt = AdamUniform([p], lr=0.05, betas=(0.9, 0.999))
    for k in range(4):
        p.grad = torch.from_numpy(a[f"g{k}"].copy()).to(DEV)
        opt.step()
        np.testing.assert_allclose(p.detach().cpu().numpy(), a[f"p{k + 1}"], rtol=0, atol=5e-7)
    assert opt.state[p]["step"] == 4


def test_adam_uniform_large_and_param_groups():
    rng = np.random.default_rng(0)
    shapes = [(100003, 3), (17,)]
    ps = [torch.nn.Parameter(torch.from_numpy(rng.normal(size=s).astype(np.float32)).to(DEV)) for s in shapes]
    oras = [oracle.AdamUniformOracle(s, lr=0.1) for s in shapes]
    cur = [p.detach().cpu().numpy().copy() for p in ps]
    opt = AdamUniform(ps)       # defaults lr=0.1, betas=(0.9, 0.999)  (optimize.py:10)
    for step in range(3):
        for i, p in enumerate(ps):
            g = rng.normal(size=shapes[i]).astype(np.float32)
            p.grad = torch.from_numpy(g).to(DEV)
            cur[i] = oras[i].step(cur[i], g)
        opt.step()
        for i, p in enumerate(ps):
            np.testing.assert_allclose(p.detach().cpu().numpy(), cur[i], rtol=0, atol=1e-6)


def test_adam_uniform_propagates_nan():
    """optimize.py:40 takes max(sqrt(m2)) with torch.max, which propagates NaN: one NaN gradient poisons every parameter and
    the divergence is visible.  (A plain fmaxf reduction would drop it.)"""
    p = torch.nn.Parameter(torch.ones(5000, 3, device=DEV))
    opt = AdamUniform([p], lr=0.1)
    g = torch.randn(5000, 3, device=DEV)
    p.grad = g.clone()
    opt.step()
    assert torch.isfinite(p).all()
    g[1234, 1] = float("nan")
    p.grad = g
    opt.step()
    assert torch.isnan(p).all()


def test_tutorial_shaped_loop_tracks_cpu_oracle():
    """Stand-in for suzanne->target (scenes and nvdiffrast are not available): source icosphere, target = displaced
    sphere with the same connectivity, loss = mean (v - v_target)^2 (the L2 image loss option of scripts/main.py:188;
    smooth, so the GPU and CPU trajectories can be compared step by step)."""
    v, f = workloads.icosphere(3)
    V = len(v)
    target = (v * (1.0 + 0.3 * np.sin(3 * v[:, :1]) * np.cos(2 * v[:, 1:2]))).astype(np.float32)
    lam = 19.0
    tv, tf = to_dev(v, f)
    M = compute_matrix(tv, tf, lam)
    u = to_differential(M, tv).clone().requires_grad_(True)
    opt = AdamUniform([u], lr=0.05)
    tgt = torch.from_numpy(target).to(DEV)
    # CPU oracle loop
    r, c, val, _ = oracle.compute_matrix(v, f, lam)
    ds = oracle.DirectSolver(r, c, val, V)
    uo = oracle.to_differential(r, c, val, V, v)
    oo = oracle.AdamUniformOracle((V, 3), lr=0.05)
    losses = []
    for step in range(200):
        x = from_differential(M, u, "Cholesky")
        loss = ((x - tgt) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
        if step < 5:
            xo = ds.solve(uo)
            go = ds.solve(2.0 * (xo - target) / (3 * V))
            uo = oo.step(uo, go.astype(np.float32))
            assert rel_l2(u.detach().cpu().numpy(), uo) < 1e-4, step
    assert losses[-1] < 0.05 * losses[0]


def test_config5_standin_2000_steps_at_52k(bunny_mesh):
    """BASELINE config 5 stand-in at paper scale: bunny subdivided x2 (V = 52,786), cotangent lambda = 19, 2000 steps of
    from_differential -> loss -> backward -> AdamUniform.  Every 100 steps the hot path is checked against the oracle ON
    THE STATE THE LOOP HAS REACHED: forward solve of the current u and backward solve of the current gradient vs the fp64
    direct solve, and the fused AdamUniform step vs the oracle step from the same moments."""
    v, f = bunny_mesh
    v, f = workloads.subdivide(*workloads.subdivide(v, f))
    v = v.astype(np.float32)
    V = len(v)
    lam = 19.0
    tv, tf = to_dev(v, f)
    M = compute_matrix(tv, tf, lam, cotan=True)
    r, c, val, _ = oracle.compute_matrix(v, f, lam, cotan=True)
    ds = oracle.DirectSolver(r, c, val, V)
    target = (v * (1.0 + 0.2 * np.sin(8 * v[:, :1]) * np.cos(6 * v[:, 1:2]))).astype(np.float32)
    tgt = torch.from_numpy(target).to(DEV)
    u = to_differential(M, tv).clone().requires_grad_(True)
    opt = AdamUniform([u], lr=0.01)
    losses, worst = [], 0.0
    for step in range(2000):
        x = from_differential(M, u, "Cholesky")
        loss = ((x - tgt) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        if step % 100 == 0:
            un = u.detach().cpu().numpy()
            e_f = rel_l2(x.detach().cpu().numpy(), ds.solve(un))
            gx = (2.0 * (x.detach() - tgt) / (3 * V)).cpu().numpy()
            e_b = rel_l2(u.grad.cpu().numpy(), ds.solve(gx))
            st = opt.state[u] if len(opt.state[u]) else None
            oo = oracle.AdamUniformOracle((V, 3), lr=0.01)
            if st is not None:
                oo.g1, oo.g2, oo.step_count = st["g1"].cpu().numpy().copy(), st["g2"].cpu().numpy().copy(), st["step"]
            want = oo.step(un.copy(), u.grad.cpu().numpy())
            opt.step()
            e_a = float(np.abs(u.detach().cpu().numpy() - want).max())
            worst = max(worst, e_f, e_b)
            assert e_f < 1e-5 and e_b < 1e-5, (step, e_f, e_b)
            assert e_a < 2e-6 * max(1.0, float(np.abs(want).max())), (step, e_a)
        else:
            opt.step()
        losses.append(float(loss.detach()) if step % 50 == 0 else None)
    ls = [l for l in losses if l is not None]
    print(f"52.8K loop: loss {ls[0]:.3e} -> {ls[-1]:.3e}, worst solve rel-L2 along the trajectory {worst:.2e}")
    assert ls[-1] < 0.05 * ls[0]
