#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 300 python profiles/spmm_probe.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({k:d[k] for k in ('env','spmm_cold_us','spmm_hot_us','solve_ms','iters','us_per_iter','phase_cycles_per_iter') if k in d}))"; }
{
run PROBE_MESH=plane
run PROBE_MESH=plane LS_PCG_PROFILE=1
run PROBE_MESH=bunny
} | tee gpurun_out/sweep15.jsonl
