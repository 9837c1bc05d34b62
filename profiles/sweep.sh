#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 300 python profiles/spmm_probe.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({k:d[k] for k in ('env','desc','spmm_cold_us','spmm_hot_us','solve_ms','iters','us_per_iter') if k in d}))"; }
{
run PROBE_MESH=plane
run PROBE_MESH=plane LS_FORCE_REORDER=1
run PROBE_MESH=bunny
run PROBE_MESH=bunny LS_PCG_MODE=graph
run PROBE_MESH=ico
run PROBE_MESH=ico LS_PCG_MODE=graph
run PROBE_MESH=plane PROBE_N=500
run PROBE_MESH=plane PROBE_N=2000
} | tee gpurun_out/sweep10.jsonl
