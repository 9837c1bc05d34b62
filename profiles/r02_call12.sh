#!/bin/bash
# bash profiles/r02_call12.sh (under gpurun): A/B of the all-reduce ring layout (-DLS_RING_SOA=1: one plane per word) on four mesh sizes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_ring_soa_ab.jsonl; : > $OUT
L=$PWD/large-steps-pytorch_b200/largesteps_b200
for rep in 1 2; do
for lib in libls_b200.so libls_b200_ringsoa.so; do
  for c in "CHK_MESH=ico CHK_PRECOND=auto" "CHK_MESH=bunny CHK_PRECOND=auto" "CHK_N=500 CHK_PRECOND=auto" "CHK_N=1000 CHK_PRECOND=jacobi"; do
    env $c CHK_DIRECT=0 CHK_REPS=100 LS_LIB_PATH=$L/$lib timeout 200 python profiles/fused_check.py 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'lib': '$lib', 'case': '$c', 'V': d['V'], 'it': d['iters'], 'ms': d['solve_ms'], 'us_it': d['us_per_iter'], 'relres': d.get('true_relres'), 'det': d['deterministic']}))" | tee -a $OUT
  done
done
done
