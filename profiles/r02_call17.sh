#!/bin/bash
# bash profiles/r02_call17.sh (under gpurun): last validation of the shipped build -- smoke, full GPU suite, bench (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider -rf 2>&1 | tail -60 > gpurun_out/r02_call5_pytest.log
tail -25 gpurun_out/r02_call5_pytest.log | cut -c1-300
echo "== bench.py"
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -3 gpurun_out/r02_bench.err | cut -c1-400
python -c "
import json
d = json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
r = d['roofline']
print(json.dumps({'value': d['value'], 'e2e': d['e2e'], 'ms': d['ms_per_step'], 'clocks': d['clocks'], 'gpu_launches': d['gpu_launches'],
                  'spmv': {k: r.get(k) for k in ('frac','us_per_launch','l2_resident_us_per_launch','traffic','with_dot_epilogue')}, 'spmv4M': r.get('spmv_4M'),
                  'solve_kernel': {k: r['solve_kernel'].get(k) for k in ('us_per_launch','us_per_iteration','phase_cycles_per_iteration','algorithmic_GBs','dram_bytes_per_launch')},
                  'cpu': d['cpu_baseline'] and {k: d['cpu_baseline'].get(k) for k in ('value','cores','quota_cores','cgroup_cpu_max','reference_cg_port','direct_solve')}, 'extra': d['extra']}, indent=1))"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench.err; cut -c1-600 gpurun_out/r02_bench_ref.json
