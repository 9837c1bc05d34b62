#!/bin/bash
# bash profiles/r02_call11.sh (under gpurun): bring-up of the cluster-resident mode (RES = 4: published rows in distributed shared
# memory) -- parity on small meshes, threshold sweep against the cooperative grid, phase cycles, sanitizer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_clres.jsonl; : > $OUT
run() { env "$@" timeout 180 python profiles/fused_check.py >> $OUT 2>gpurun_out/r02_clres.err || { echo "FAILED: $*"; tail -5 gpurun_out/r02_clres.err; }; }
echo "== parity (direct solve) in the new mode"
run CHK_MESH=ico CHK_PRECOND=auto
run CHK_MESH=ico CHK_PRECOND=auto LS_PCG_PROFILE=1
run CHK_MESH=ico CHK_PRECOND=auto LS_PCG_CLRES=0
run CHK_MESH=ico CHK_LEVEL=5 CHK_PRECOND=auto
run CHK_MESH=ico CHK_LEVEL=5 CHK_PRECOND=auto LS_PCG_CLRES=0
run CHK_MESH=ico CHK_LEVEL=5 CHK_PRECOND=auto LS_PCG_PATTERN=0
run CHK_N=70 CHK_ALPHA=0.999 CHK_PRECOND=auto
run CHK_N=70 CHK_ALPHA=0.999 CHK_PRECOND=auto LS_PCG_CLRES=0
echo "== threshold sweep: cluster-resident (forced) vs cooperative grid, preconditioner auto"
for n in 40 50 64 90 110 128 160 181; do
  run CHK_N=$n CHK_PRECOND=auto CHK_DIRECT=0 LS_PCG_CLRES=100000
  run CHK_N=$n CHK_PRECOND=auto CHK_DIRECT=0 LS_PCG_CLRES=0
done
run CHK_N=64 CHK_PRECOND=auto CHK_DIRECT=0 LS_PCG_CLRES=100000 LS_PCG_SMALLCTA=0
run CHK_N=110 CHK_PRECOND=auto CHK_DIRECT=0 LS_PCG_CLRES=100000 LS_PCG_PROFILE=1
python - <<'PY'
import json
for l in open('gpurun_out/r02_clres.jsonl'):
    d = json.loads(l)
    e = {k: v for k, v in d['env'].items()}
    print(json.dumps({'env': e, 'V': d['V'], 'grid': d['desc'].get('grid'), 'res': d['desc'].get('residency'), 'thr': d['desc'].get('threads'), 'pre': d['desc'].get('precond'),
                      'it': d['iters'], 'st': d['status'], 'ms': d['solve_ms'], 'us_it': d['us_per_iter'], 'err': [d.get('err_fwd'), d.get('err_bwd'), d.get('true_relres')],
                      'det': d['deterministic'], 'cyc': d.get('phase_cycles_per_iter')}))
PY
echo "== pytest (solver tests)"
timeout 1500 python -m pytest tests/test_gpu_pcg.py tests/test_gpu_adam_loop.py tests/test_gpu_remesh.py -m gpu -q -x --timeout 900 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -15 | cut -c1-300
echo "== sanitizer (quick)"
SAN_QUICK=1 timeout 1500 bash profiles/sanitizer.sh 2>&1 | tail -30
cp gpurun_out/sanitizer.log gpurun_out/r02_sanitizer_clres.log
