#!/usr/bin/env python
"""BASELINE config 5 stand-in (nvdiffrast and the scenes are unavailable): Tutorial-shaped optimisation loop
   from_differential -> loss -> backward (second solve) -> AdamUniform.step,  2000 steps,
source = icosphere / subdivided bunny, target = displaced copy with the same connectivity, loss = mean squared distance.
Prints it/s (two solves + one fused optimiser step per iteration; no renderer in the loop)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "large-steps-pytorch_b200"))
import numpy as np
import torch
from largesteps_b200 import workloads
from largesteps.geometry import compute_matrix
from largesteps.parameterize import to_differential, from_differential
from largesteps.optimize import AdamUniform

dev = "cuda:0"
out = []
for name in ("icosphere4", "bunny2"):
    if name == "icosphere4":
        v, f = workloads.icosphere(4)
    else:
        d = np.load(os.path.join(ROOT, "tests", "golden", "bunny_mesh.npz"))
        v, f = workloads.subdivide(*workloads.subdivide(d["verts"], d["faces"].astype(np.int64)))
        v = v.astype(np.float32)
    target = torch.from_numpy((v * (1.0 + 0.3 * np.sin(3 * v[:, :1]) * np.cos(2 * v[:, 1:2]))).astype(np.float32)).to(dev)
    tv, tf = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
    M = compute_matrix(tv, tf, lambda_=19)
    u = to_differential(M, tv).clone().requires_grad_(True)
    opt = AdamUniform([u], lr=0.01)
    steps = 2000

    def run(n):
        last = None
        for _ in range(n):
            x = from_differential(M, u, 'Cholesky')
            loss = ((x - target) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            last = loss
        return last

    l0 = float(((from_differential(M, u.detach(), 'Cholesky') - target) ** 2).mean())
    run(20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last = run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out.append({"mesh": name, "V": int(v.shape[0]), "steps": steps, "it_per_s": steps / dt, "ms_per_it": 1e3 * dt / steps,
                "loss_start": l0, "loss_end": float(last)})
print(json.dumps(out))
