"""GPU parity: y = M x (csrc/ls_spmm_kernel.cuh, TMA-staged CSR SpMM) vs scipy on the same matrix."""
import numpy as np
import pytest
import torch

import oracle
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import to_differential, spmm
from gpu_util import DEV, to_dev, rel_l2, fan_mesh, coo_np

pytestmark = pytest.mark.gpu
TOL = 2e-6   # fp32 SpMM vs fp64 scipy, relative L2


def scipy_of(M):
    idx, val = coo_np(M)
    return oracle.coo_to_scipy(idx[0], idx[1], val, M.shape[0])


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 9])
def test_spmm_columns(bunny_mesh, k):
    v, f = bunny_mesh
    M = compute_matrix(*to_dev(v, f), 19.0, cotan=True)
    x = np.random.default_rng(k).normal(size=(len(v), k)).astype(np.float32)
    y = to_differential(M, torch.from_numpy(x).to(DEV))
    assert y.shape == (len(v), k) and y.dtype == torch.float32
    assert rel_l2(y.cpu().numpy(), scipy_of(M) @ x.astype(np.float64)) < TOL


def test_spmm_vector_and_golden(golden_assembly, golden_solve):
    g, s = golden_assembly, golden_solve
    for mesh, kw in (("ico2", dict(lambda_=10.0)), ("bunny", dict(lambda_=19.0, cotan=True))):
        M = compute_matrix(*to_dev(g[f"{mesh}.verts"], g[f"{mesh}.faces"]), **kw)
        u = to_differential(M, torch.from_numpy(s[f"{mesh}.v"]).to(DEV))
        assert rel_l2(u.cpu().numpy(), s[f"{mesh}.u"]) < 1e-5          # the reference's own torch `M @ v`
        y1 = spmm(M, torch.from_numpy(s[f"{mesh}.v"][:, 0].copy()).to(DEV))
        assert y1.dim() == 1 and rel_l2(y1.cpu().numpy(), u[:, 0].cpu().numpy()) < 1e-6


def test_spmm_matches_torch_sparse_on_device():
    v, f = workloads.plane(200, seed=1)
    M = compute_matrix(*to_dev(v, f), 1.0, alpha=0.95)
    x = torch.randn(M.shape[0], 3, device=DEV)
    assert rel_l2(to_differential(M, x).cpu().numpy(), (M @ x).cpu().numpy()) < 1e-6


def test_long_rows_and_ragged_blocks():
    # hub row (6001 nnz) is longer than a stage: exercises the direct path and the shrinking-block search
    v, f = fan_mesh(6000)
    M = compute_matrix(*to_dev(v, f), 1.0, alpha=0.5)
    x = np.random.default_rng(0).normal(size=(len(v), 3)).astype(np.float32)
    y = to_differential(M, torch.from_numpy(x).to(DEV))
    assert rel_l2(y.cpu().numpy(), scipy_of(M) @ x.astype(np.float64)) < TOL
    # several medium hubs: blocks that do not fit a stage but whose rows do
    rng = np.random.default_rng(1)
    V = 5000
    rows = np.concatenate([np.repeat(np.arange(0, V, 50), 700), np.arange(V)])
    cols = np.concatenate([rng.integers(0, V, size=len(rows) - V), np.arange(V)])
    vals = rng.normal(size=len(rows)).astype(np.float32)
    A = torch.sparse_coo_tensor(torch.from_numpy(np.stack([rows, cols])).to(DEV), torch.from_numpy(vals).to(DEV), (V, V)).coalesce()
    y = spmm(A, torch.from_numpy(x[:V]).to(DEV))
    assert rel_l2(y.cpu().numpy(), scipy_of(A) @ x[:V].astype(np.float64)) < 5e-6


def test_empty_rows_and_tiny():
    # general CSR with empty rows (not a system matrix): rows 1 and 3 empty
    A = torch.sparse_coo_tensor(torch.tensor([[0, 0, 2, 4], [0, 4, 2, 1]], device=DEV),
                                torch.tensor([1., 2., 3., 4.], device=DEV), (5, 5)).coalesce()
    x = torch.arange(10, dtype=torch.float32, device=DEV).view(5, 2)
    assert torch.equal(spmm(A, x), (A @ x))
    # 1x1
    B = torch.sparse_coo_tensor(torch.tensor([[0], [0]], device=DEV), torch.tensor([2.5], device=DEV), (1, 1)).coalesce()
    assert spmm(B, torch.tensor([[2.0, 4.0]], device=DEV)).cpu().tolist() == [[5.0, 10.0]]


def test_shuffled_vertex_order(bunny_mesh):
    v, f = bunny_mesh
    v, f = workloads.subdivide(v, f)
    v, f = workloads.shuffle_vertices(v.astype(np.float32), f, seed=5)
    M = compute_matrix(*to_dev(v, f), 19.0, cotan=True)
    x = np.random.default_rng(2).normal(size=(len(v), 3)).astype(np.float32)
    assert rel_l2(to_differential(M, torch.from_numpy(x).to(DEV)).cpu().numpy(), scipy_of(M) @ x.astype(np.float64)) < TOL


def test_to_differential_is_differentiable():
    v, f = workloads.icosphere(2)
    M = compute_matrix(*to_dev(v, f), 10.0)
    x = torch.randn(len(v), 3, device=DEV, requires_grad=True)
    g = torch.randn(len(v), 3, device=DEV)
    (to_differential(M, x) * g).sum().backward()
    assert rel_l2(x.grad.cpu().numpy(), (M.t() @ g).cpu().numpy()) < 1e-6


def test_full_size_linearity_and_constant_vector():
    """V = 1e6 (BASELINE config 3): M 1 = (1-alpha) 1 exactly-ish, and linearity."""
    v, f = workloads.plane(1000, seed=0)
    M = compute_matrix(*to_dev(v, f), 1.0, alpha=0.95)
    V = M.shape[0]
    one = torch.ones(V, 3, device=DEV)
    y = to_differential(M, one)
    assert float((y - 0.05).abs().max()) < 5e-6
    a = torch.randn(V, 3, device=DEV)
    b = torch.randn(V, 3, device=DEV)
    lhs = to_differential(M, 2.0 * a + b)
    rhs = 2.0 * to_differential(M, a) + to_differential(M, b)
    assert rel_l2(lhs.cpu().numpy(), rhs.cpu().numpy()) < 1e-6
    assert rel_l2(lhs.cpu().numpy(), (M @ (2.0 * a + b)).cpu().numpy()) < 1e-6
