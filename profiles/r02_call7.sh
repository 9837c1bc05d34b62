#!/bin/bash
# bash profiles/r02_call7.sh (under gpurun): A/B of the bf16 published rows (ZH) against the fp32 rows, then the full GPU suite on the new default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=$PWD/large-steps-pytorch_b200/largesteps_b200
OUT=gpurun_out/r02_call7_check.jsonl
: > $OUT
chk() { env "$@" timeout 240 python profiles/fused_check.py 2>gpurun_out/chk.err | tail -1 | tee -a $OUT | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); e = d['env']; e.pop('LS_LIB_PATH', None)
print({k: d.get(k) for k in ('env','iters','restarts','solve_ms','us_per_iter','true_relres','err_fwd','err_bwd','deterministic','phase_cycles_per_iter')})"; tail -2 gpurun_out/chk.err | cut -c1-300; }
for lib in libls_b200.so libls_b200_zf32.so; do
  echo "== $lib"
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 LS_PCG_PROFILE=1
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 LS_PCG_PATTERN=0
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=500 CHK_DIRECT=0
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=300
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=300 CHK_ALPHA=0.999
  chk LS_LIB_PATH=$D/$lib CHK_MESH=bunny
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=2000 CHK_DIRECT=0 CHK_REPS=5
done
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider -rf 2>&1 | tail -60 > gpurun_out/r02_call7_pytest.log
tail -25 gpurun_out/r02_call7_pytest.log | cut -c1-300
