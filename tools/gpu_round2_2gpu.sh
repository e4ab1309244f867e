#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02_2gpu}
nvidia-smi -L | head -4
SECONDS=0
timeout 500 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -s --timeout=450 -p no:cacheprovider > gpurun_out/pytest_dp_nccl_$TAG.log 2>&1
echo "dp test: $(tail -1 gpurun_out/pytest_dp_nccl_$TAG.log) [${SECONDS}s]"
grep -E "Error|assert" gpurun_out/pytest_dp_nccl_$TAG.log | head -5 | cut -c1-300
SECONDS=0
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$? [${SECONDS}s]"
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ('value','ms_per_step','n_gpus')}); print(j.get('train_step')); print(j['config'].get('parallelism'))
except Exception as e: print("ERR", e)
PY
tail -5 gpurun_out/bench_$TAG.err | cut -c1-300
