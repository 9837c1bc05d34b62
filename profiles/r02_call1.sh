#!/bin/bash
# bash profiles/r02_call1.sh (under gpurun): first run of the fused two-synchronisation solver + TMA-staged SELL SpMM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_call1_check.jsonl
: > $OUT
chk() { env "$@" timeout 240 python profiles/fused_check.py 2>gpurun_out/chk.err | tail -1 | tee -a $OUT | cut -c1-600; tail -3 gpurun_out/chk.err | cut -c1-300; }
echo "== correctness + timing, fused vs classic"
chk CHK_MESH=ico
chk CHK_MESH=ico LS_PCG_ALGO=classic
chk CHK_MESH=plane CHK_N=64
chk CHK_MESH=bunny LS_PCG_PROFILE=1
chk CHK_MESH=bunny LS_PCG_CLUSTER=0
chk CHK_MESH=bunny LS_PCG_ALGO=classic
chk CHK_MESH=plane CHK_N=300 LS_PCG_PROFILE=1
chk CHK_MESH=plane CHK_N=300 CHK_ALPHA=0.999
chk CHK_MESH=plane CHK_N=300 CHK_ALPHA=0.999 LS_PCG_ALGO=classic
chk CHK_MESH=plane CHK_N=500 CHK_DIRECT=0 LS_PCG_PROFILE=1
chk CHK_MESH=plane CHK_N=500 CHK_DIRECT=0 LS_PCG_ALGO=classic
chk CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0
chk CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 LS_PCG_PROFILE=1
chk CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 LS_PCG_PATTERN=0
chk CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0 LS_PCG_ALGO=classic
chk CHK_MESH=plane CHK_N=2000 CHK_DIRECT=0 CHK_REPS=5
chk CHK_MESH=plane CHK_N=2000 CHK_DIRECT=0 CHK_REPS=5 LS_PCG_ALGO=classic
echo "== stand-alone SELL SpMM: register prefetch (0) vs TMA-staged variants"
for v in 0 1 2 3 4 11; do
  LS_SELL_TMA=$v LS_PCG_MODE=graph timeout 300 python profiles/spmm_probe.py 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(json.dumps({'LS_SELL_TMA': $v, **{k: d[k] for k in ('spmm_cold_us','spmm_hot_us','spmm_cold_GBs','iter3_cold_us','solve_ms','iters') if k in d}}))" | tee -a gpurun_out/r02_call1_sell.jsonl
done
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02_call1_pytest.log
tail -25 gpurun_out/r02_call1_pytest.log
echo "== compute-sanitizer (fused solver)"
SAN_QUICK=1 bash profiles/sanitizer.sh > /dev/null 2>&1
cp gpurun_out/sanitizer.log gpurun_out/r02_call1_sanitizer.log; cat gpurun_out/sanitizer.log
