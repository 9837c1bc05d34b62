#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 300 python profiles/spmm_probe.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({k:d[k] for k in ('env','spmm_cold_us','spmm_hot_us','solve_ms','us_per_iter') if k in d}))"; }
D=$PWD/large-steps-pytorch_b200/largesteps_b200
{
run PROBE_MESH=plane
run PROBE_MESH=plane LS_LIB_PATH=$D/libls_b200_nopf.so
run PROBE_MESH=plane
run PROBE_MESH=plane LS_LIB_PATH=$D/libls_b200_nopf.so
} | tee gpurun_out/sweep18.jsonl
