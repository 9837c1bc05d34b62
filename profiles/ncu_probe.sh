#!/bin/bash
# bash profiles/ncu_probe.sh <tag> <kernel-regex> <skip> [ENV=VAL ...]  (under gpurun): --set full capture of one launch in spmm_probe.py
TAG=$1; KREG=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
env "$@" ncu --set full --clock-control none --import-source on -k regex:${KREG} -s ${SKIP} -c 1 -o gpurun_out/${TAG} -f \
    python profiles/spmm_probe.py > gpurun_out/${TAG}.log 2>&1
