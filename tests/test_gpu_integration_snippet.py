"""GPU: the binding INTEGRATION.md tells a maintainer of the reference to add (raw ctypes on the C ABI, no
largesteps_b200 wrappers) works as written and meets the parity bar."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import oracle
from conftest import ROOT
from largesteps_b200 import workloads, _native as N
from largesteps_b200.geometry import compute_matrix
from gpu_util import DEV, to_dev, rel_l2

pytestmark = pytest.mark.gpu


def snippet_source():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n# largesteps/solvers.py  \(addition\).*?```", md, flags=re.S).group(0)
    code = block[len("```python\n"):-3]
    code = code.replace('ctypes.CDLL("libls_b200.so")', f'ctypes.CDLL(r"{N.LIB_PATH}")')
    code = code.split("# largesteps/parameterize.py:50-55")[0]
    return code


def test_integration_md_binding_runs_and_matches_direct_solve():
    class Solver:                       # the reference's base class (solvers.py:6-24)
        def __init__(self, M):
            pass

    ns = {"Solver": Solver, "ctypes": ctypes, "torch": torch}
    exec(compile(snippet_source(), "INTEGRATION.md", "exec"), ns)
    v, f = workloads.icosphere(4)
    kw = dict(lambda_=10.0)
    M = compute_matrix(*to_dev(v, f), **kw)
    solver = ns["B200Solver"](M)
    r, c, val, V = oracle.compute_matrix(v, f, **kw)
    b = np.random.default_rng(0).normal(size=(V, 3)).astype(np.float32)
    x = solver.solve(torch.from_numpy(b).to(DEV))
    torch.cuda.synchronize()
    assert rel_l2(x.cpu().numpy(), oracle.DirectSolver(r, c, val, V).solve(b)) < 1e-5
    del solver


def test_pure_c_host_roundtrip(tmp_path):
    """examples/c_host/roundtrip.c: assembly -> to_differential -> solve -> AdamUniform through the C ABI from a plain C
    program (no Python, no torch in the process); the program itself checks ||v - verts|| / ||verts|| <= 1e-5."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h"):
        pytest.skip("gcc or the CUDA headers are not available")
    out = str(tmp_path / "roundtrip")
    libdir = os.path.dirname(N.LIB_PATH)
    cudart_dirs = ["/usr/local/cuda/lib64"]
    r = subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-I", "/usr/local/cuda/include", os.path.join(ROOT, "examples", "c_host", "roundtrip.c"), "-o", out,
                        "-L", libdir, "-l:" + os.path.basename(N.LIB_PATH), "-Wl,-rpath," + libdir,
                        "-L", cudart_dirs[0], "-Wl,-rpath," + cudart_dirs[0], "-lcudart", "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for n in ("64", "700"):
        r = subprocess.run([out, n], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "c-abi roundtrip ok" in r.stdout
