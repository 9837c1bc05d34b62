#!/bin/bash
# bash profiles/pattern_ab.sh   (under gpurun, ~4 min): everything needed to decide whether LS_PCG_PATTERN becomes the default
#   1. default vs pattern-only at V = 1e6, twice each, same box: solves/s, ms, per-phase cycles -> gpurun_out/pattern_ab.jsonl
#   2. the PCG parity tests with the opt-in set in the environment                            -> gpurun_out/pattern_tests.log
#      (tests/test_gpu_pcg.py asserts sell_engine == 1 in one place: expect that one to flag engine 2)
#   3. mid-size and small meshes (bunny x2, icosphere, plane500, plane2000) both ways         -> gpurun_out/pattern_sizes.jsonl
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/pattern_ab.jsonl
: > gpurun_out/pattern_sizes.jsonl
one() {  # $1 = 0/1 pattern, $2 = workload, $3 = steps, $4 = output file
  LS_PCG_PATTERN=$1 timeout 300 python bench.py --workload "$2" --steps "$3" --warmup 5 --no-cpu-baseline --no-spmv-4m 2>/dev/null \
    | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print(json.dumps({'pattern': $1, 'workload': '$2', 'value': round(d['value'], 1), 'ms': round(d['ms_per_step'], 4),
                  'e2e': round(d['e2e']['value'], 1), 'phases': r.get('phase_cycles_per_iteration')}))" | tee -a "$4"
}
for rep in 1 2; do
  for pat in 0 1; do one $pat plane1000 50 gpurun_out/pattern_ab.jsonl; done
done
LS_PCG_PATTERN=1 timeout 600 python -m pytest tests/test_gpu_pcg.py -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -15 > gpurun_out/pattern_tests.log
cat gpurun_out/pattern_tests.log
for wl in bunny icosphere plane500 plane2000; do
  for pat in 0 1; do one $pat $wl 30 gpurun_out/pattern_sizes.jsonl; done
done
