#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" python profiles/spmm_probe.py 2>&1 | tail -1; }
{
run LS_SPMM_ENGINE=sell
run LS_SPMM_ENGINE=csr
run LS_SPMM_ENGINE=sell PROBE_REORDER=0
run LS_SPMM_ENGINE=csr PROBE_REORDER=0
run LS_SPMM_ENGINE=sell LS_VEC_MODE=0
} | tee gpurun_out/sweep5.jsonl
