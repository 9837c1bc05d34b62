#!/usr/bin/env python
"""GPU probe used while tuning: times the in-solver SpMM (HBM-cold by rotation, and L2-resident) and a full solve
for the current LS_SPMM_* environment.  One JSON line.  Not a bench value source -- bench.py is."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "large-steps-pytorch_b200"))
import numpy as np
import torch
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import to_differential
from largesteps_b200.solvers import PCGSolver

n = int(os.environ.get("PROBE_N", "1000"))
dev = "cuda:0"
mesh = os.environ.get("PROBE_MESH", "plane")
if mesh == "ico":
    v, f = workloads.icosphere(4)
    kw = dict(lambda_=10.0)
elif mesh == "bunny":
    d = np.load(os.path.join(ROOT, "tests", "golden", "bunny_mesh.npz"))
    v, f = workloads.subdivide(*workloads.subdivide(d["verts"], d["faces"].astype(np.int64)))
    v = v.astype(np.float32)
    kw = dict(lambda_=19.0, cotan=True)
else:
    v, f = workloads.plane(n, seed=0)
    kw = dict(lambda_=1.0, alpha=0.95)
tv, tf = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
M = compute_matrix(tv, tf, **kw)
hs = [PCGSolver(M, reorder=os.environ.get('PROBE_REORDER', '1') == '1') for _ in range(int(os.environ.get('PROBE_HANDLES', '4')))]
u = to_differential(M, tv + 0.01 * torch.randn_like(tv))


from largesteps_b200.solvers import bench_kernels


def timeit(which, handles, reps=400):
    bench_kernels(handles, which, 8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bench_kernels(handles, which, reps)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


out = {"env": {k: v for k, v in os.environ.items() if k.startswith(("LS_", "PROBE_"))}, "n": n, "desc": hs[0].describe()}
names = ["spmm", "update", "pupdate", "iter3", "spmv_nodot"]
for w in range(5):
    out[names[w] + "_cold_us"] = round(timeit(w, hs), 2)
    out[names[w] + "_hot_us"] = round(timeit(w, hs[:1]), 2)
B = hs[0].spmm_bytes(3)
out["spmm_cold_GBs"] = round(B / out["spmm_cold_us"] / 1e3, 1)
hs[0].solve(u)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    hs[0].solve(u)
e1.record()
torch.cuda.synchronize()
out["solve_ms"] = round(e0.elapsed_time(e1) / 20, 3)
out["iters"] = hs[0].iterations
out["us_per_iter"] = round(1e3 * out["solve_ms"] / out["iters"], 2)
if os.environ.get("LS_PCG_PROFILE"):
    pc = hs[0].phase_cycles(per_cta=True)
    itn = max(pc["iterations"], 1)
    out["phase_cycles_per_iter"] = {k: round(v / itn) for k, v in pc.items() if k not in ("_", "iterations", "per_cta")}
    import numpy as _np
    t = _np.array(pc["per_cta"], dtype=_np.float64) / itn
    for j, nm in enumerate(["spmm", "reduce1", "update", "reduce2", "pupdate", "barrier3"]):
        out["cta_" + nm] = {"min": round(t[:, j].min()), "med": round(float(_np.median(t[:, j]))), "max": round(t[:, j].max())}
    order = _np.argsort(t[:, 0])
    out["slowest_spmm_ctas"] = [(int(c), int(t[c, 6] * itn), round(t[c, 0])) for c in order[-6:]]
    out["fastest_spmm_ctas"] = [(int(c), int(t[c, 6] * itn), round(t[c, 0])) for c in order[:6]]
print(json.dumps(out))
