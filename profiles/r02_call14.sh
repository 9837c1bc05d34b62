#!/bin/bash
# bash profiles/r02_call14.sh (under gpurun): phase A instruction diet -- mixed-precision adds for the bf16 rows (LS_FHADD) and
# float4 rows for x / p at RES = 1 (LS_XP4): A/B of the four builds on four mesh sizes, then the solver tests on the default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/${OUTNAME:-r02_phaseA_diet_ab}.jsonl; : > $OUT
L=$PWD/large-steps-pytorch_b200/largesteps_b200
for rep in 1 2; do
for lib in ${LIBS:-libls_b200_base.so libls_b200_fhadd.so libls_b200_xp4.so libls_b200.so}; do
  for c in "CHK_MESH=ico CHK_PRECOND=auto" "CHK_MESH=bunny CHK_PRECOND=auto" "CHK_N=500 CHK_PRECOND=auto" "CHK_N=1000 CHK_PRECOND=jacobi" "CHK_N=1000 CHK_PRECOND=jacobi LS_PCG_PATTERN=0"; do
    env $c CHK_DIRECT=0 CHK_REPS=100 LS_LIB_PATH=$L/$lib timeout 200 python profiles/fused_check.py 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'lib': '$lib', 'case': '$c', 'V': d['V'], 'it': d['iters'], 'ms': d['solve_ms'], 'us_it': d['us_per_iter'], 'relres': d.get('true_relres'), 'det': d['deterministic']}))" | tee -a $OUT
  done
done
done
echo "== phase cycles, default build, 1M"
CHK_N=1000 CHK_DIRECT=0 LS_PCG_PROFILE=1 timeout 200 python profiles/fused_check.py 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({'ms': d['solve_ms'], 'cyc': d.get('phase_cycles_per_iter'), 'cta': d.get('cta_minmedmax')}))" | tee -a $OUT
echo "== parity vs direct (default build)"
for c in "CHK_N=300" "CHK_N=300 CHK_ALPHA=0.999" "CHK_N=700" "CHK_MESH=bunny" "CHK_N=700 LS_PCG_RES=1" "CHK_N=300 LS_PCG_RES=1 LS_PCG_PATTERN=0"; do
  env $c timeout 300 python profiles/fused_check.py 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({'case': '$c', 'res': d['desc'].get('residency'), 'it': d['iters'], 'err': [d.get('err_fwd'), d.get('err_bwd')], 'ms': d['solve_ms'], 'det': d['deterministic']}))" | tee -a $OUT
done
echo "== pytest (solver tests, default build)"
timeout 1500 python -m pytest tests/test_gpu_pcg.py tests/test_gpu_pattern.py tests/test_gpu_adam_loop.py -m gpu -q -x --timeout 900 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -6 | cut -c1-300
