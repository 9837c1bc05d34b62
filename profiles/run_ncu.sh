#!/bin/bash
# Run under gpurun (1 GPU):  bash profiles/run_ncu.sh <tag>
# Produces gpurun_out/<tag>_launches.csv (every launch with its device time) and gpurun_out/<tag>_prof.ncu-rep
# (--set full capture of the three CG-iteration kernels).  Numbers printed by bench.py under ncu are NOT bench values.
TAG=${1:-r01}
mkdir -p gpurun_out
export LS_BENCH_WORKLOAD=${LS_BENCH_WORKLOAD:-plane1000}
ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 640 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spmm_tma|k_update|k_pupdate" -s 30 -c 6 \
    -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out/
