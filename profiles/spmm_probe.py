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
v, f = workloads.plane(n, seed=0)
tv, tf = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
M = compute_matrix(tv, tf, 1.0, alpha=0.95)
hs = [PCGSolver(M) for _ in range(4)]
u = to_differential(M, tv + 0.01 * torch.randn_like(tv))


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


L = 400
cold = timeit(lambda i=0: hs[i % 4].bench_spmm(3, 1), L)
hot = timeit(lambda i=0: hs[0].bench_spmm(3, 1), L)
solve = timeit(lambda i=0: hs[0].solve(u), 20)
it = hs[0].iterations
B = hs[0].spmm_bytes(3)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("LS_")}, "n": n,
                  "spmm_cold_us": round(cold, 2), "spmm_hot_us": round(hot, 2), "cold_GBs": round(B / cold / 1e3, 1),
                  "solve_ms": round(solve / 1e3, 3), "iters": it, "us_per_iter": round(solve / it, 2)}))
