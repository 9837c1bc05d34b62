#!/bin/bash
# bash profiles/r02_call2.sh (under gpurun): second round-2 GPU run -- everything-in-shared-memory single CTA (RES = 3), byte
# cuts at 1M, Chebyshev preconditioner, loop glue / remesh tests, SELL TMA variants, racecheck details, bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_call2_check.jsonl
: > $OUT
chk() { env "$@" timeout 240 python profiles/fused_check.py 2>gpurun_out/chk.err | tail -1 | tee -a $OUT | cut -c1-900; tail -2 gpurun_out/chk.err | cut -c1-300; }
echo "== fused solver: sizes"
chk CHK_MESH=ico LS_PCG_PROFILE=1
chk CHK_MESH=ico
chk CHK_MESH=plane CHK_N=40
chk CHK_MESH=plane CHK_N=64
chk CHK_MESH=plane CHK_N=64 LS_PCG_CLUSTER=16
chk CHK_MESH=bunny
chk CHK_MESH=bunny CHK_PRECOND=chebyshev
chk CHK_MESH=plane CHK_N=500 CHK_DIRECT=0 LS_PCG_PROFILE=1
chk CHK_MESH=plane CHK_N=500 CHK_DIRECT=0
chk CHK_MESH=plane CHK_N=500 CHK_ALPHA=0.999 CHK_DIRECT=0
chk CHK_MESH=plane CHK_N=500 CHK_ALPHA=0.999 CHK_DIRECT=0 CHK_PRECOND=chebyshev
chk CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0
chk CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 LS_PCG_PROFILE=1
chk CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 CHK_PRECOND=chebyshev
chk CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 LS_PCG_REFINE=0
echo "== stand-alone SELL SpMM variants (with and without the dot epilogue)"
for v in 1 3 5 6; do
  LS_SELL_TMA=$v LS_PCG_MODE=graph timeout 300 python profiles/spmm_probe.py 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(json.dumps({'LS_SELL_TMA': $v, **{k: d[k] for k in ('spmm_cold_us','spmm_hot_us','spmv_nodot_cold_us','spmv_nodot_hot_us','iter3_cold_us') if k in d}}))" | tee -a gpurun_out/r02_call2_sell.jsonl
done
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider -rf 2>&1 | tail -60 > gpurun_out/r02_call2_pytest.log
tail -30 gpurun_out/r02_call2_pytest.log
echo "== racecheck details"
cat > /tmp/san2.py <<'PY'
import os, sys
sys.path.insert(0, "large-steps-pytorch_b200"); sys.path.insert(0, ".")
import numpy as np, torch
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import to_differential
from largesteps_b200.solvers import PCGSolver
v, f = workloads.plane(int(os.environ.get("SAN_N", "70")), seed=0)
tv, tf = torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()
M = compute_matrix(tv, tf, 1.0, alpha=0.95)
s = PCGSolver(M, maxit=6, precond=os.environ.get("SAN_PRECOND", "jacobi"))
import warnings; warnings.simplefilter("ignore")
x = s.solve(to_differential(M, tv)); torch.cuda.synchronize(); print("san2 ok", s.describe())
PY
for mode in "LS_PCG_CLUSTER=16 SAN_N=70" "SAN_N=40" "SAN_N=70 SAN_PRECOND=chebyshev" "SAN_N=70 LS_PCG_CLUSTER=8 SAN_PRECOND=chebyshev"; do
  echo "=== racecheck $mode"
  env $mode timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python /tmp/san2.py 2>&1 | grep -v "^$" | head -45 | cut -c1-260
done > gpurun_out/r02_call2_racecheck.log 2>&1
head -120 gpurun_out/r02_call2_racecheck.log
echo "== bench.py"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02_call2_bench.json 2> gpurun_out/r02_call2_bench.err; tail -3 gpurun_out/r02_call2_bench.err | cut -c1-400
python -c "
import json
d = json.loads(open('gpurun_out/r02_call2_bench.json').read().strip().splitlines()[-1])
r = d['roofline']
print(json.dumps({'value': d['value'], 'e2e': d['e2e']['value'], 'ms': d['ms_per_step'], 'spmv_frac': r['frac'], 'spmv_us': r['us_per_launch'], 'spmv4M': r.get('spmv_4M', {}).get('frac'),
                  'solve_kernel': {k: r['solve_kernel'].get(k) for k in ('us_per_launch','us_per_iteration','phase_cycles_per_iteration')}, 'cpu': d['cpu_baseline'] and {k: d['cpu_baseline'][k] for k in ('value','cores','quota_cores','cgroup_cpu_max')}, 'extra': d['extra']}, indent=1))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_call2_bench_ref.json 2>> gpurun_out/r02_call2_bench.err; cut -c1-1500 gpurun_out/r02_call2_bench_ref.json
