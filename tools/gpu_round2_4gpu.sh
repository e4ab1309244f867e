#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02_4gpu}
N=${2:-4}
nvidia-smi -L | head -8
SECONDS=0
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$? [${SECONDS}s]"
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ('value','ms_per_step','n_gpus')}, 'e2e', j['e2e']['value']); print(j.get('train_step')); print(j.get('clocks'))
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/bench_$TAG.err | cut -c1-300
SECONDS=0
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "reference arm under torchrun rc=$? [${SECONDS}s]"; tail -1 gpurun_out/bench_ref_$TAG.json | cut -c1-200
