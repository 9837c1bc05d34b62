#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 300 python profiles/spmm_probe.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({k:d[k] for k in ('env','solve_ms','iters','us_per_iter') if k in d}))"; }
{
run PROBE_MESH=bunny
run PROBE_MESH=bunny LS_PCG_SMALLCTA=0
run PROBE_MESH=plane PROBE_N=300
run PROBE_MESH=plane PROBE_N=300 LS_PCG_SMALLCTA=0
run PROBE_MESH=plane
} | tee gpurun_out/sweep16.jsonl
