#!/usr/bin/env python
"""GPU check used while bringing up the fused solver (csrc/ls_pcg_fused.cuh): parity vs the CPU direct solve and solve
times for one mesh under the current LS_* environment.  One JSON line per call.  Not a bench value source."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "large-steps-pytorch_b200"))
import numpy as np
import torch
import oracle
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import to_differential
from largesteps_b200.solvers import PCGSolver

mesh = os.environ.get("CHK_MESH", "plane")
n = int(os.environ.get("CHK_N", "300"))
alpha = float(os.environ.get("CHK_ALPHA", "0.95"))
direct = os.environ.get("CHK_DIRECT", "1") == "1"
if mesh == "ico":
    v, f = workloads.icosphere(int(os.environ.get("CHK_LEVEL", "4")))
    kw = dict(lambda_=10.0)
elif mesh == "bunny":
    d = np.load(os.path.join(ROOT, "tests", "golden", "bunny_mesh.npz"))
    v, f = workloads.subdivide(*workloads.subdivide(d["verts"], d["faces"].astype(np.int64)))
    v = v.astype(np.float32)
    kw = dict(lambda_=19.0, cotan=True)
else:
    v, f = workloads.plane(n, seed=0)
    kw = dict(lambda_=1.0, alpha=alpha)
dev = "cuda:0"
tv, tf = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
M = compute_matrix(tv, tf, **kw)
s = PCGSolver(M, check=False, precond=os.environ.get("CHK_PRECOND", "jacobi"))
out = {"env": {k: v_ for k, v_ in os.environ.items() if k.startswith(("LS_", "CHK_"))}, "V": int(M.shape[0]), "desc": s.describe()}
r, c, val, V = oracle.compute_matrix(v, f, **kw)
A = oracle.coo_to_scipy(r, c, val, V)
_, b, g = workloads.rhs_recipe(lambda x: A @ x, v)
tb, tg = torch.from_numpy(b).to(dev), torch.from_numpy(g).to(dev)
x = s.solve(tb)
torch.cuda.synchronize()
out["iters"], out["status"], out["restarts"], out["relres"] = s.iterations, s.status, s.restarts, [float("%.2e" % q) for q in s.relres[:3]]
xg = s.solve(tg, backward=True)
out["iters_bwd"], out["restarts_bwd"] = s.iterations, s.restarts
if direct:
    ds = oracle.DirectSolver(r, c, val, V)
    rel = lambda a_, b_: float(np.linalg.norm(a_.astype(np.float64) - b_) / np.linalg.norm(b_))
    out["err_fwd"] = float("%.3e" % rel(x.cpu().numpy(), ds.solve(b)))
    out["err_bwd"] = float("%.3e" % rel(xg.cpu().numpy(), ds.solve(g)))
else:
    res = (M @ x - tb).norm(dim=0) / tb.norm(dim=0)
    out["true_relres"] = float("%.3e" % float(res.max()))
x2 = s.solve(tb)
out["deterministic"] = bool(torch.equal(x, x2))
reps = int(os.environ.get("CHK_REPS", "20"))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    s.solve(tb)
e1.record()
torch.cuda.synchronize()
out["solve_ms"] = round(e0.elapsed_time(e1) / reps, 4)
out["us_per_iter"] = round(1e3 * out["solve_ms"] / max(out["iters"], 1), 3)
if os.environ.get("LS_PCG_PROFILE"):
    pc = s.phase_cycles(per_cta=True)
    itn = max(pc["iterations"], 1)
    out["phase_cycles_per_iter"] = {k: round(v_ / itn) for k, v_ in pc.items() if k not in ("_", "_0", "iterations", "per_cta")}
    tab = np.array(pc["per_cta"], dtype=np.float64) / itn
    out["cta_minmedmax"] = {nm: [round(tab[:, j].min()), round(float(np.median(tab[:, j]))), round(tab[:, j].max())]
                            for j, nm in enumerate(list(pc.keys())[:4])}
print(json.dumps(out))
