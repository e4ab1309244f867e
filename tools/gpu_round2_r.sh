#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02r}
for f in test_gpu_detector test_gpu_vs_reference; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
grep -h "AssertionError\|Error" gpurun_out/pytest_test_gpu_*_$TAG.log | head -10 | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 --no-reference-gpu --no-descriptor > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['value']); print(j.get('train_step')); print(j.get('train_step_tf32_backward')); print(j['clocks'])
PY
tail -3 gpurun_out/bench_$TAG.err
