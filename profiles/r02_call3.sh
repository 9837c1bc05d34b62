#!/bin/bash
# bash profiles/r02_call3.sh (under gpurun): meshops test details, SELL prefetch A/B, CTA-size A/B of the fused solver, ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=$PWD/large-steps-pytorch_b200/largesteps_b200
echo "== meshops tests"
timeout 600 python -m pytest tests/test_gpu_meshops.py tests/test_gpu_remesh.py -m gpu -q -x --timeout 300 -p no:cacheprovider -s 2>&1 | tail -40 | cut -c1-400
echo "== SELL SpMV: variants x L2 prefetch"
for cfg in "3 1024" "3 0" "7 1024" "1 1024" "2 1024"; do
  set -- $cfg
  LS_SELL_TMA=$1 LS_SELL_PF=$2 LS_PCG_MODE=graph timeout 300 python profiles/spmm_probe.py 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(json.dumps({'LS_SELL_TMA': $1, 'LS_SELL_PF': $2, **{k: d[k] for k in ('spmm_cold_us','spmm_hot_us','spmv_nodot_cold_us','spmv_nodot_hot_us','iter3_cold_us') if k in d}}))" | tee -a gpurun_out/r02_call3_sell.jsonl
done
echo "== fused solver: CTA size A/B (768 / 640 / 512 threads)"
OUT=gpurun_out/r02_call3_check.jsonl
: > $OUT
chk() { env "$@" timeout 240 python profiles/fused_check.py 2>gpurun_out/chk.err | tail -1 | tee -a $OUT | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print({k: d.get(k) for k in ('env','iters','solve_ms','us_per_iter','phase_cycles_per_iter','true_relres','err_fwd')})"; tail -2 gpurun_out/chk.err | cut -c1-300; }
for lib in libls_b200.so libls_b200_pt640.so libls_b200_pt512.so; do
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 LS_PCG_PROFILE=1
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=500 CHK_DIRECT=0 LS_PCG_SMALLCTA=0
  chk LS_LIB_PATH=$D/$lib CHK_MESH=bunny LS_PCG_SMALLCTA=0
done
chk CHK_MESH=bunny
chk CHK_MESH=plane CHK_N=64 LS_PCG_CLUSTER=16
echo "== racecheck cluster 16 (after the start barrier)"
cat > /tmp/san2.py <<'PY'
import os, sys
sys.path.insert(0, "large-steps-pytorch_b200"); sys.path.insert(0, ".")
import numpy as np, torch, warnings
from largesteps_b200 import workloads
from largesteps_b200.geometry import compute_matrix
from largesteps_b200.parameterize import to_differential
from largesteps_b200.solvers import PCGSolver
v, f = workloads.plane(70, seed=0)
tv, tf = torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()
M = compute_matrix(tv, tf, 1.0, alpha=0.95)
s = PCGSolver(M, maxit=6)
warnings.simplefilter("ignore")
x = s.solve(to_differential(M, tv)); torch.cuda.synchronize(); print("san2 ok", s.describe())
PY
LS_PCG_CLUSTER=16 timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python /tmp/san2.py 2>&1 | grep -v "^$" | head -30 | cut -c1-260 | tee gpurun_out/r02_call3_racecheck.log
echo "== ncu: stand-alone SpMV (TMA-staged, no dot) and the fused solve at 1M"
LS_PCG_MODE=graph PROBE_HANDLES=4 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:spmm_sell_tma_kernel<3, false" -s 12 -c 2 -o gpurun_out/r02_sell_tma_1M -f \
    python profiles/spmm_probe.py > gpurun_out/r02_sell_tma_1M.log 2>&1
PROBE_HANDLES=1 ncu --set full --clock-control none -k regex:pcg_fused -s 2 -c 1 -o gpurun_out/r02_fused_1M -f \
    python profiles/spmm_probe.py > gpurun_out/r02_fused_1M.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spmv-4m --no-config4 > gpurun_out/r02_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
