#!/bin/bash
# bash profiles/r02_call0.sh  (under gpurun, 1 GPU): round-2 opening measurements on the round-1 build
#   1. pattern-only copy vs general copy at every mesh size, with per-phase / per-CTA cycle tables
#   2. ncu --set full of the shipped 768-thread spmm_sell_kernel at V = 1e6 and of one persistent solve (both formats)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_call0_probe.jsonl
: > $OUT
probe() { env "$@" timeout 300 python profiles/spmm_probe.py 2>/dev/null | tail -1 >> $OUT; }
for pat in 0 1; do
  probe LS_PCG_PATTERN=$pat PROBE_MESH=plane
  probe LS_PCG_PATTERN=$pat PROBE_MESH=plane LS_PCG_PROFILE=1
  probe LS_PCG_PATTERN=$pat PROBE_MESH=plane PROBE_N=500 PROBE_HANDLES=1 LS_PCG_PROFILE=1
  probe LS_PCG_PATTERN=$pat PROBE_MESH=ico PROBE_HANDLES=1 LS_PCG_PROFILE=1
done
probe PROBE_MESH=bunny PROBE_HANDLES=1 LS_PCG_PROFILE=1
probe LS_PCG_PATTERN=1 PROBE_MESH=plane PROBE_N=2000 PROBE_HANDLES=1
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: d[k] for k in ('env','spmm_cold_us','spmm_hot_us','solve_ms','iters','us_per_iter','phase_cycles_per_iter') if k in d})"
# ncu: the stand-alone SELL kernel at 1M (first cold-rotation launches) and one whole solve per format
ncu --set full --clock-control none --import-source on -k regex:spmm_sell -s 12 -c 2 -o gpurun_out/r02_sell_1M -f \
    python profiles/spmm_probe.py > gpurun_out/r02_sell_1M.log 2>&1
for pat in 0 1; do
  LS_PCG_PATTERN=$pat PROBE_HANDLES=1 ncu --set full --clock-control none -k regex:pcg_persistent -s 2 -c 1 -o gpurun_out/r02_persist_pat$pat -f \
    python profiles/spmm_probe.py > gpurun_out/r02_persist_pat$pat.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
