"""GPU parity: on-device assembly (csrc/ls_assemble.cu through the C ABI) vs the reference's own outputs
(tests/golden/assembly.npz) and vs the oracle at larger sizes.  Integer structure: bit-exact.  Uniform values:
bit-exact.  Cotangent values: 2e-6 of max|M| (the reference's own fp32 summation order is unspecified)."""
import numpy as np
import pytest
import torch

import oracle
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix, laplacian_uniform, laplacian_cot, csr_of
from gpu_util import DEV, to_dev, coo_np, fan_mesh

pytestmark = pytest.mark.gpu

MESHES = ["tet", "quad", "ico2", "bunny", "odd"]
CASES = {"uni_l10": dict(lambda_=10.0), "uni_a095": dict(lambda_=1.0, alpha=0.95),
         "cot_l19": dict(lambda_=19.0, cotan=True), "cot_a09": dict(lambda_=1.0, alpha=0.9, cotan=True)}


def check_against(M, idx, gv, exact):
    gi, gval = coo_np(M)
    assert M.dtype == torch.float32 and M.indices().dtype == torch.int64
    assert gi.shape == idx.shape, (gi.shape, idx.shape)
    assert (gi == idx).all()
    if exact:
        assert (gval == gv).all()
    else:
        assert np.abs(gval - gv).max() <= 2e-6 * np.abs(gv).max()


@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("mesh", MESHES)
@pytest.mark.parametrize("case", list(CASES))
def test_compute_matrix_vs_reference_golden(golden_assembly, mesh, case, idx_dtype):
    g = golden_assembly
    v, f = to_dev(g[f"{mesh}.verts"], g[f"{mesh}.faces"], idx_dtype)
    M = compute_matrix(v, f, **CASES[case])
    check_against(M, g[f"{mesh}.{case}.idx"], g[f"{mesh}.{case}.val"], exact=case.startswith("uni"))


@pytest.mark.parametrize("mesh", MESHES)
def test_laplacians_vs_reference_golden(golden_assembly, mesh):
    g = golden_assembly
    v, f = to_dev(g[f"{mesh}.verts"], g[f"{mesh}.faces"])
    Lc = laplacian_cot(v, f)
    check_against(Lc, g[f"{mesh}.Lcot.idx"], g[f"{mesh}.Lcot.val"], exact=False)
    Lu = laplacian_uniform(v, f)
    gi, gval = coo_np(Lu)
    ridx, rval = g[f"{mesh}.Luni.idx"], g[f"{mesh}.Luni.val"]
    if mesh == "odd":
        # documented difference: an isolated vertex gets an explicit 0 diagonal here, no entry in the reference
        keep = ~((gi[0] == gi[1]) & (gval == 0))
        gi, gval = gi[:, keep], gval[keep]
    assert (gi == ridx).all() and (gval == rval).all()


def test_csr_matches_coo(golden_assembly):
    g = golden_assembly
    v, f = to_dev(g["bunny.verts"], g["bunny.faces"])
    M = compute_matrix(v, f, 19.0, cotan=True)
    rowptr, col, val = csr_of(M)
    idx, vals = coo_np(M)
    rp = rowptr.cpu().numpy()
    assert rp[0] == 0 and rp[-1] == idx.shape[1]
    assert (np.repeat(np.arange(len(rp) - 1), np.diff(rp)) == idx[0]).all()
    assert (col.cpu().numpy() == idx[1]).all() and (val.cpu().numpy() == vals).all()
    assert col.dtype == torch.int32 and rowptr.dtype == torch.int32


def test_foreign_coo_to_csr():
    """A matrix NOT built by compute_matrix (torch's own ops on the GPU) goes through ls_coo_to_csr."""
    v, f = workloads.plane(40)
    r, c, val, V = oracle.compute_matrix(v, f, 5.0)
    M = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])).to(DEV), torch.from_numpy(val).to(DEV), (V, V)).coalesce()
    rowptr, col, vv = csr_of(M)
    rp = rowptr.cpu().numpy()
    assert (np.repeat(np.arange(V), np.diff(rp)) == r).all() and (col.cpu().numpy() == c).all()
    assert (vv.cpu().numpy() == val).all()
    # uncoalesced input is coalesced first
    Mu = torch.sparse_coo_tensor(torch.tensor([[1, 0, 1], [1, 0, 1]], device=DEV), torch.tensor([1., 2., 3.], device=DEV), (2, 2))
    rowptr, col, vv = csr_of(Mu)
    assert rowptr.cpu().tolist() == [0, 1, 2] and vv.cpu().tolist() == [2.0, 4.0]


def test_edge_cases():
    # no faces at all: M = shift * I
    v = torch.rand(5, 3, device=DEV)
    M = compute_matrix(v, torch.zeros((0, 3), dtype=torch.int64, device=DEV), 3.0)
    idx, val = coo_np(M)
    assert (idx[0] == np.arange(5)).all() and (idx[1] == np.arange(5)).all() and (val == 1).all()
    # degenerate faces (repeated vertex) and a duplicated face: same as the oracle's coalesce semantics
    vv = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32)
    ff = np.array([[0, 1, 2], [0, 1, 2], [0, 0, 3], [1, 3, 3], [2, 3, 0]], dtype=np.int64)
    for kw in (dict(lambda_=2.0), dict(lambda_=1.0, alpha=0.5), dict(lambda_=2.0, cotan=True)):
        r, c, val, V = oracle.compute_matrix(vv, ff, **kw)
        M = compute_matrix(*to_dev(vv, ff), **kw)
        gi, gval = coo_np(M)
        assert (gi[0] == r).all() and (gi[1] == c).all()
        assert np.abs(gval - val).max() <= 1e-5 * max(np.abs(val).max(), 1.0)
    # out-of-range index
    bad = torch.tensor([[0, 1, 9]], device=DEV)
    with pytest.raises(IndexError):
        compute_matrix(torch.rand(4, 3, device=DEV), bad, 1.0)
    with pytest.raises(IndexError):
        compute_matrix(torch.rand(4, 3, device=DEV), torch.tensor([[0, -1, 2]], device=DEV), 1.0)
    with pytest.raises(ValueError):
        compute_matrix(torch.rand(4, 3, device=DEV), torch.zeros((2, 4), dtype=torch.int64, device=DEV), 1.0)


def test_high_valence_hub():
    v, f = fan_mesh(6000)
    for kw in (dict(lambda_=1.0, alpha=0.9), dict(lambda_=3.0, cotan=True)):
        r, c, val, V = oracle.compute_matrix(v, f, **kw)
        M = compute_matrix(*to_dev(v, f), **kw)
        gi, gval = coo_np(M)
        assert (gi[0] == r).all() and (gi[1] == c).all()
        assert np.abs(gval - val).max() <= 2e-5 * np.abs(val).max()


@pytest.mark.parametrize("kw", [dict(lambda_=1.0, alpha=0.95), dict(lambda_=19.0, cotan=True)])
def test_mid_size_vs_oracle(bunny_mesh, kw):
    v, f = bunny_mesh
    v, f = workloads.subdivide(v, f)
    v = v.astype(np.float32)
    v, f = workloads.shuffle_vertices(v, f, seed=3)       # arbitrary (bad) vertex order
    r, c, val, V = oracle.compute_matrix(v, f, **kw)
    M = compute_matrix(*to_dev(v, f), **kw)
    gi, gval = coo_np(M)
    assert (gi[0] == r).all() and (gi[1] == c).all()
    if kw.get("cotan"):
        assert np.abs(gval - val).max() <= 2e-6 * np.abs(val).max()
    else:
        assert (gval == val).all()


def test_full_size_invariants():
    """BASELINE config 3 at full size (V = 1e6): size-independent properties."""
    n = 1000
    v, f = workloads.plane(n, seed=0)
    tv, tf = to_dev(v, f)
    M = compute_matrix(tv, tf, 1.0, alpha=0.95)
    V = n * n
    assert M.shape == (V, V) and M._nnz() == 6992002 and M.is_coalesced()
    idx = M.indices()
    val = M.values()
    key = idx[0] * V + idx[1]
    assert bool((key[1:] > key[:-1]).all())                     # strictly row-major sorted, no duplicates
    # symmetric: the transposed key set with the same values is the same multiset
    tkey = idx[1] * V + idx[0]
    order = torch.argsort(tkey)
    assert bool((tkey[order] == key).all()) and bool((val[order] == val).all())
    # uniform: off-diagonals all -alpha, diagonal = (1-alpha) + alpha * deg, L 1 = 0  =>  M 1 = (1-alpha) 1
    off = idx[0] != idx[1]
    assert bool((val[off] == np.float32(-0.95)).all())
    deg = torch.zeros(V, device=DEV).index_add_(0, idx[0][off], torch.ones(int(off.sum()), device=DEV))
    diag = val[~off]
    expect = (torch.tensor(np.float32(1 - 0.95), device=DEV) + torch.tensor(np.float32(0.95), device=DEV) * deg)
    assert bool((diag == expect).all())
    assert int(deg.min()) == 2 and int(deg.max()) == 6
    # cotangent at full size: row sums of L are ~0, matrix symmetric to rounding
    L = laplacian_cot(tv, tf)
    rs = torch.zeros(V, device=DEV, dtype=torch.float64).index_add_(0, L.indices()[0], L.values().double())
    assert float(rs.abs().max()) <= 1e-4 * float(L.values().abs().max())
