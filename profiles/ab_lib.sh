#!/bin/bash
# A/B two builds of the native library on the same box: profiles/ab_lib.sh libA.so libB.so  -> gpurun_out/ab_lib.jsonl
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/ab_lib.jsonl
for rep in 1 2; do
  for lib in "$@"; do
    LS_LIB_PATH="$PWD/large-steps-pytorch_b200/largesteps_b200/$lib" timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-spmv-4m 2>/dev/null \
      | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print(json.dumps({'lib': '$lib', 'value': round(d['value'], 1), 'ms': round(d['ms_per_step'], 4), 'e2e': round(d['e2e']['value'], 1), 'phases': r.get('phase_cycles_per_iteration')}))" | tee -a gpurun_out/ab_lib.jsonl
  done
done
