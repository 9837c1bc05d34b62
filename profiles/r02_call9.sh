#!/bin/bash
# bash profiles/r02_call9.sh (under gpurun): evidence for the shipped build -- ncu of the fused solve (bf16 rows) and the SpMV at 1M,
# launch list of the bench command, compute-sanitizer on the bf16-row paths
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PROBE_HANDLES=1 ncu --set full --clock-control none -k regex:pcg_fused -s 2 -c 1 -o gpurun_out/r02_fused_1M -f \
    python profiles/spmm_probe.py > gpurun_out/r02_fused_1M.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 500 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spmv-4m > gpurun_out/r02_launches_final.log 2>&1
ls -la gpurun_out/r02_fused_1M.ncu-rep gpurun_out/r02_launches_final.csv
SAN_QUICK=1 bash profiles/sanitizer.sh > /dev/null 2>&1; cp gpurun_out/sanitizer.log gpurun_out/r02_call9_sanitizer.log; cat gpurun_out/sanitizer.log
