#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02s}
for f in test_gpu_ops test_gpu_detector; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
grep -h "AssertionError\|Error" gpurun_out/pytest_test_gpu_*_$TAG.log | head -10 | cut -c1-300
SECONDS=0
timeout 1200 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
echo "bench default rc=$? [${SECONDS}s]"
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step','steps','warmup','gpu_launches')}, 'e2e', j['e2e']['value']); print(j.get('train_step')); print(j.get('train_step_tf32_backward'))
print(j.get('roofline',{}).get('kernel'), j.get('roofline',{}).get('frac'), j.get('clocks'))
d=j.get('descriptor',{}); print({k:(d.get(k) or {}).get('ms_median', (d.get(k) or {}).get('ours_ms', (d.get(k) or {}).get('ms'))) for k in ('ball_group_fused','index_max_op','descriptor_forward_eval')})
print(j.get('reference_gpu',{}).get('tf32_off')); print(j.get('reference_gpu',{}).get('torch_default'))
PY
tail -3 gpurun_out/bench_$TAG.err
