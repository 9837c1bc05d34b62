#!/bin/bash
# bash profiles/r02_call16.sh (under gpurun): does 60 KB of L1 instead of 28 KB help the V = 1e6 solve?  The grid instantiations no
# longer carry the 4 KB cluster exchange area, so without the pattern diagonal in shared memory (LS_PCG_DP=0) the kernel fits the
# 196 KB carve-out.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_l1_carveout_ab.jsonl; : > $OUT
for rep in 1 2; do
for dp in "LS_X=1" "LS_PCG_DP=0"; do
  for c in "CHK_N=1000 CHK_PRECOND=jacobi" "CHK_N=1000 CHK_PRECOND=jacobi LS_PCG_PATTERN=0" "CHK_N=900 CHK_PRECOND=jacobi"; do
    env $c $dp CHK_DIRECT=0 CHK_REPS=100 timeout 200 python profiles/fused_check.py 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'dp': '$dp', 'case': '$c', 'V': d['V'], 'res': d['desc'].get('residency'), 'it': d['iters'], 'ms': d['solve_ms'], 'relres': d.get('true_relres'), 'det': d['deterministic']}))" | tee -a $OUT
  done
done
done
