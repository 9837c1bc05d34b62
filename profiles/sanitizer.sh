#!/bin/bash
# bash profiles/sanitizer.sh  (under gpurun): compute-sanitizer memcheck + racecheck over a small solve in every mode
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import os, sys
sys.path.insert(0, "large-steps-pytorch_b200"); sys.path.insert(0, ".")
import numpy as np, torch, oracle
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import to_differential, from_differential
from largesteps_b200.solvers import PCGSolver, ConjugateGradientSolver
from largesteps_b200.optimize import AdamUniform
v, f = workloads.plane(70, seed=0)
tv, tf = torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()
for kw in (dict(lambda_=1.0, alpha=0.95), dict(lambda_=19.0, cotan=True)):
    M = compute_matrix(tv, tf, **kw)
    u = to_differential(M, tv).clone().requires_grad_(True)
    x = from_differential(M, u)
    x.sum().backward()
    s = ConjugateGradientSolver(M); s.solve(u.detach()); s.solve(u.detach() * 1.01)
    r, c, val, V = oracle.compute_matrix(v, f, **kw)
    xd = oracle.DirectSolver(r, c, val, V).solve(u.detach().cpu().numpy())
    err = np.linalg.norm(x.detach().cpu().numpy() - xd) / np.linalg.norm(xd)
    assert err < 1e-5, err
    opt = AdamUniform([u], lr=0.01); opt.step()
torch.cuda.synchronize(); print("sanitizer case ok", {k: v for k, v in os.environ.items() if k.startswith("LS_")})
PY
for tool in memcheck racecheck; do
  # SAN_QUICK=1: fused solver only (one cluster, cooperative grid at three residency levels)
  for mode in "LS_X=1" "LS_PCG_CLUSTER=0" "LS_PCG_CLUSTER=0 LS_PCG_RES=1" "LS_PCG_CLUSTER=0 LS_PCG_RES=0 LS_PCG_PATTERN=0" ${SAN_QUICK:+--} "LS_PCG_ALGO=classic" "LS_PCG_MODE=graph" "LS_PCG_MODE=graph LS_SPMM_ENGINE=csr"; do
    if [ "$mode" = "--" ]; then break; fi
    echo "=== $tool $mode"
    env $mode timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san_case.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer case ok|Error|hazard|access at" | head -12
  done
done 2>&1 | tee gpurun_out/sanitizer.log
