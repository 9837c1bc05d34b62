#!/bin/bash
# bash profiles/r02_call4.sh (under gpurun): Chebyshev steps through the optimised gather -- timings, full test suite, sanitizer, SpMV ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_call4_check.jsonl
: > $OUT
chk() { env "$@" timeout 240 python profiles/fused_check.py 2>gpurun_out/chk.err | tail -1 | tee -a $OUT | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print({k: d.get(k) for k in ('env','iters','restarts','solve_ms','us_per_iter','true_relres','err_fwd','err_bwd','deterministic')}, {k: d['desc'].get(k) for k in ('grid','cluster','residency','threads','sell_engine')})"; tail -2 gpurun_out/chk.err | cut -c1-300; }
echo "== Jacobi vs Chebyshev"
for mesh in "CHK_MESH=ico" "CHK_MESH=plane CHK_N=64" "CHK_MESH=bunny" "CHK_MESH=plane CHK_N=300" "CHK_MESH=plane CHK_N=500 CHK_DIRECT=0" "CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0" "CHK_MESH=plane CHK_N=500 CHK_ALPHA=0.999 CHK_DIRECT=0"; do
  chk $mesh
  chk $mesh CHK_PRECOND=chebyshev
  chk $mesh CHK_PRECOND=chebyshev LS_PCG_CHEB_M=3
done
chk CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 CHK_PRECOND=chebyshev LS_PCG_CHEB_M=6
chk CHK_MESH=bunny CHK_PRECOND=chebyshev LS_PCG_CHEB_M=6
chk CHK_MESH=plane CHK_N=2000 CHK_DIRECT=0 CHK_REPS=5 CHK_PRECOND=chebyshev
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider -rf 2>&1 | tail -80 > gpurun_out/r02_call4_pytest.log
tail -30 gpurun_out/r02_call4_pytest.log | cut -c1-300
echo "== compute-sanitizer: Chebyshev paths"
cat > /tmp/san3.py <<'PY'
import os, sys
sys.path.insert(0, "large-steps-pytorch_b200"); sys.path.insert(0, ".")
import numpy as np, torch, warnings
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import to_differential
from largesteps_b200.solvers import PCGSolver
warnings.simplefilter("ignore")
for kw in (dict(lambda_=1.0, alpha=0.95), dict(lambda_=19.0, cotan=True)):
    v, f = workloads.plane(70, seed=0)
    tv, tf = torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()
    M = compute_matrix(tv, tf, **kw)
    s = PCGSolver(M, maxit=5, precond="chebyshev", warm_start=True)
    x = s.solve(to_differential(M, tv)); x = s.solve(to_differential(M, tv) * 1.01)
torch.cuda.synchronize(); print("san3 ok", s.describe())
PY
for tool in memcheck racecheck; do
  for mode in "LS_X=1" "LS_PCG_RES=1" "LS_PCG_RES=0" "LS_PCG_CLUSTER=8"; do
    echo "=== $tool $mode"
    env $mode timeout 400 compute-sanitizer --tool $tool python /tmp/san3.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|san3 ok|Error|hazard|Invalid" | head -6 | cut -c1-240
  done
done 2>&1 | tee gpurun_out/r02_call4_sanitizer.log
echo "== ncu: stand-alone SpMV (TMA-staged, no dot) at 1M"
LS_PCG_MODE=graph PROBE_HANDLES=4 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:spmm_sell_tma_kernel<\(int\)3, \(bool\)0' -s 12 -c 2 -o gpurun_out/r02_sell_tma_1M -f \
    python profiles/spmm_probe.py > gpurun_out/r02_sell_tma_1M.log 2>&1
ls -la gpurun_out/r02_sell_tma_1M.ncu-rep; tail -2 gpurun_out/r02_sell_tma_1M.log | cut -c1-200
