#!/bin/bash
# bash profiles/r02_call6.sh (under gpurun --gpus 2): the multi-rank path of bench.py (torchrun, NCCL) and the 2-GPU test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 \
  > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -3 gpurun_out/r02_bench_n2.err | cut -c1-300
python -c "
import json
d = json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ('value','n_gpus','ms_per_step','e2e','gpu_launches','clocks')}, indent=1)); print(json.dumps(d['extra']['config4'], indent=1))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_pcg.py -m gpu -q -k "two_gpus or config4" -p no:cacheprovider 2>&1 | tail -3
