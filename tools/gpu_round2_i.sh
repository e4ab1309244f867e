#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02i}
for f in test_gpu_ops test_gpu_detector test_gpu_vs_reference test_dropin_imports; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
grep -h "ablation .* worst\|AssertionError" gpurun_out/pytest_test_gpu_detector_$TAG.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step')}); print(j.get('train_step')); print(j.get('train_step_tf32_backward'))
d=j.get('descriptor',{}); print({k:d.get(k) for k in ('ball_group_fused','index_max_op','descriptor_forward_eval')})
print(j.get('reference_gpu',{}).get('tf32_off'))
PY
tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ball_launches_$TAG.csv python tools/ncu_step.py ballonly > /dev/null 2>&1
grep -E "bx_|index_max" gpurun_out/ball_launches_$TAG.csv | tail -6 | awk -F'","' '{print substr($5,1,40), $NF}'
