#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 300 python profiles/spmm_probe.py 2>&1 | tail -1; }
{
run LS_PCG_PROFILE=1 PROBE_REORDER=0
run PROBE_REORDER=0
run PROBE_REORDER=0 LS_PCG_FASTRED=0
run PROBE_REORDER=1
} | tee gpurun_out/sweep9.jsonl
