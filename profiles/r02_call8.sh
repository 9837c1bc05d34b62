#!/bin/bash
# bash profiles/r02_call8.sh (under gpurun): A/B of the polling variant of the publishing all-reduce, then bench.py on the default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=$PWD/large-steps-pytorch_b200/largesteps_b200
OUT=gpurun_out/r02_call8_check.jsonl
: > $OUT
chk() { env "$@" timeout 240 python profiles/fused_check.py 2>gpurun_out/chk.err | tail -1 | tee -a $OUT | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); e = d['env']; lib = e.pop('LS_LIB_PATH', '')[-24:]
print(lib, {k: d.get(k) for k in ('env','iters','solve_ms','us_per_iter','true_relres','err_fwd','deterministic')})"; tail -2 gpurun_out/chk.err | cut -c1-300; }
for rep in 1 2; do
for lib in libls_b200.so libls_b200_pollfence.so; do
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=1000 CHK_DIRECT=0
  chk LS_LIB_PATH=$D/$lib CHK_MESH=plane CHK_N=500 CHK_DIRECT=0 CHK_PRECOND=jacobi
  chk LS_LIB_PATH=$D/$lib CHK_MESH=bunny
  chk LS_LIB_PATH=$D/$lib CHK_MESH=bunny CHK_PRECOND=chebyshev
done
done
echo "== bench.py"
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -3 gpurun_out/r02_bench.err | cut -c1-400
python -c "
import json
d = json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
r = d['roofline']
print(json.dumps({'value': d['value'], 'e2e': d['e2e']['value'], 'e2e_serial': d['e2e']['serial_one_stream_value'], 'ms': d['ms_per_step'], 'clocks': d['clocks'],
                  'spmv_frac': r['frac'], 'solve_kernel': {k: r['solve_kernel'].get(k) for k in ('us_per_launch','us_per_iteration','phase_cycles_per_iteration')},
                  'cpu': d['cpu_baseline'] and d['cpu_baseline']['value'], 'extra': {k: v for k, v in d['extra'].items() if k != 'checksums'}}, indent=1))"
