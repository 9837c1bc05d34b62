"""GPU parity: the per-step loop glue (csrc/ls_glue.cu through largesteps_b200.meshops) against outputs AND gradients of
the unmodified reference (scripts/geometry.py, scripts/main.py:176-180,192-195) run on the CPU by
tests/golden/make_golden_glue.py.  Forward values vs the reference's own float32 run; gradients vs its float64 run."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from largesteps_b200 import meshops
from largesteps_b200.geometry import laplacian_uniform
from gpu_util import DEV, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def glue():
    return np.load(os.path.join(GOLDEN, "glue.npz"))


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


@pytest.mark.parametrize("mesh", ["ico2", "bunny"])
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_loop_glue_forward_and_gradients(glue, mesh, idx_dtype):
    g = glue
    v_src, f_src = t(g[f"{mesh}.v_src"]), t(g[f"{mesh}.f_src"])
    # remove_duplicates: same unique vertices (sorted rows), same faces, same inverse map as the reference
    v_unique, f_unique, dup = meshops.remove_duplicates(v_src, f_src)
    assert torch.equal(v_unique.cpu(), torch.from_numpy(g[f"{mesh}.v_unique"]))
    assert torch.equal(f_unique.cpu(), torch.from_numpy(g[f"{mesh}.f_unique"]))
    assert torch.equal(dup.cpu(), torch.from_numpy(g[f"{mesh}.dup"]))
    np.testing.assert_allclose(float(meshops.average_edge_length(v_unique, f_unique)), float(g[f"{mesh}.avg_edge"]), rtol=2e-6)
    faces = f_unique.to(idx_dtype)
    dupi = dup.to(idx_dtype)
    x = v_unique.clone().requires_grad_(True)
    v_opt = meshops.gather_rows(x, dupi)                         # scripts/main.py:176
    fn = meshops.compute_face_normals(x, faces)                  # :178
    n_unique = meshops.compute_vertex_normals(x, faces, fn)      # :179
    n_opt = meshops.gather_rows(n_unique, dupi)                  # :180
    assert torch.equal(v_opt, x.detach()[dup])
    assert fn.shape == (3, faces.shape[0]) and n_unique.shape == x.shape
    np.testing.assert_allclose(fn.detach().cpu().numpy(), g[f"{mesh}.f32.face_normals"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(n_unique.detach().cpu().numpy(), g[f"{mesh}.f32.vertex_normals"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(n_opt.detach().cpu().numpy(), g[f"{mesh}.f32.n_opt"], rtol=0, atol=3e-6)
    loss = (v_opt * t(g[f"{mesh}.W1"])).sum() + (n_opt * t(g[f"{mesh}.W2"])).sum() + (fn * t(g[f"{mesh}.W3"])).sum()
    loss.backward()
    assert abs(float(loss.detach()) - float(g[f"{mesh}.f64.loss"])) <= 1e-4 * max(1.0, abs(float(g[f"{mesh}.f64.loss"])))
    err = rel_l2(x.grad.cpu().numpy(), g[f"{mesh}.f64.grad"])
    ref_err = rel_l2(g[f"{mesh}.f32.grad"], g[f"{mesh}.f64.grad"])       # what the reference's own float32 run achieves
    print(f"{mesh}: gradient rel-L2 vs the reference's float64 run {err:.2e} (the reference's float32 run: {ref_err:.2e})")
    assert err < max(5e-6, 20 * ref_err), (err, ref_err)
    # bit-reproducible: the same graph evaluated twice gives the same bits (no atomics on the per-step path; the order in which
    # autograd adds the four contributions to x.grad depends on the graph, so the graph must be the same)
    def run():
        xx = v_unique.clone().requires_grad_(True)
        vo = meshops.gather_rows(xx, dupi)
        f_n = meshops.compute_face_normals(xx, faces)
        no = meshops.gather_rows(meshops.compute_vertex_normals(xx, faces, f_n), dupi)
        ((vo * t(g[f"{mesh}.W1"])).sum() + (no * t(g[f"{mesh}.W2"])).sum() + (f_n * t(g[f"{mesh}.W3"])).sum()).backward()
        return no.detach(), xx.grad
    n_a, g_a = run()
    n_b, g_b = run()
    assert torch.equal(n_a, n_b) and torch.equal(g_a, g_b)
    assert torch.equal(n_a, n_opt.detach()) and torch.equal(g_a, x.grad)


@pytest.mark.parametrize("mesh", ["ico2", "bunny"])
def test_laplacian_regularizer(glue, mesh):
    g = glue
    v_unique, f_unique = t(g[f"{mesh}.v_unique"]), t(g[f"{mesh}.f_unique"])
    L = laplacian_uniform(v_unique, f_unique)
    for bil, name in ((True, "bilap"), (False, "lap")):
        y = v_unique.clone().requires_grad_(True)
        reg = meshops.laplacian_regularizer(L, y, bilaplacian=bil)
        reg.backward()
        np.testing.assert_allclose(float(reg), float(g[f"{mesh}.f64.{name}"]), rtol=2e-5)
        assert rel_l2(y.grad.cpu().numpy(), g[f"{mesh}.f64.{name}_grad"]) < 2e-5


def test_glue_edge_cases():
    # empty index vector, index out of range, hub vertex (long incidence list), wrong shapes / dtypes
    v = torch.randn(10, 3, device=DEV)
    assert meshops.gather_rows(v, torch.zeros(0, dtype=torch.int64, device=DEV)).shape == (0, 3)
    bad = v.clone().requires_grad_(True)
    out = meshops.gather_rows(bad, torch.tensor([0, 3, 3], device=DEV))
    out.sum().backward()
    assert bad.grad[3, 0].item() == 2.0 and bad.grad[1, 0].item() == 0.0
    from gpu_util import fan_mesh
    hv, hf = fan_mesh(3000)
    tv, tf = t(hv), t(hf)
    x = tv.clone().requires_grad_(True)
    n = meshops.compute_vertex_normals(x, tf, meshops.compute_face_normals(x, tf))
    n.sum().backward()
    assert torch.isfinite(n).all() and torch.isfinite(x.grad).all()
    assert abs(float(n[0].norm()) - 1.0) < 1e-5
    with pytest.raises(IndexError):
        meshops.compute_vertex_normals(tv, torch.tensor([[0, 1, 5000]], device=DEV), torch.zeros(3, 1, device=DEV))
    with pytest.raises(ValueError):
        meshops.compute_face_normals(tv[:, :2], tf)
    with pytest.raises(RuntimeError):
        meshops.compute_face_normals(tv.cpu(), tf)
