#!/bin/bash
# bash profiles/sweep.sh  (under gpurun): SpMM / vector-kernel configuration sweep
mkdir -p gpurun_out
run() { env "$@" python profiles/spmm_probe.py 2>/dev/null | tail -1; }
{
run LS_SPMM_UNROLL=4 LS_SPMM_HINT=0 LS_VEC_MODE=0
run LS_SPMM_UNROLL=8 LS_SPMM_HINT=0 LS_VEC_MODE=0
run LS_SPMM_UNROLL=8 LS_SPMM_HINT=1 LS_VEC_MODE=0
run LS_SPMM_UNROLL=8 LS_SPMM_HINT=0 LS_VEC_MODE=1
run LS_SPMM_UNROLL=8 LS_SPMM_HINT=1 LS_VEC_MODE=1
run LS_SPMM_UNROLL=8 LS_SPMM_HINT=2 LS_VEC_MODE=1
run LS_SPMM_UNROLL=8 LS_SPMM_HINT=0 LS_VEC_MODE=1 LS_SPMM_STAGES=3
run LS_SPMM_UNROLL=8 LS_SPMM_HINT=0 LS_VEC_MODE=1 LS_SPMM_CAPMUL=12
run LS_SPMM_UNROLL=4 LS_SPMM_HINT=1 LS_VEC_MODE=1 LS_SPMM_STAGES=3
} | tee gpurun_out/sweep.jsonl
