#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02f}
for f in test_gpu_ops test_gpu_detector test_gpu_vs_reference test_dropin_imports; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step')}); print(j.get('train_step')); print(j.get('train_step_tf32_backward'))
d=j.get('descriptor',{}); print({k:d.get(k) for k in ('ball_group_fused','index_max_op','descriptor_forward_eval')})
PY
tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ball_launches_$TAG.csv python tools/ncu_step.py ballonly > /dev/null 2>&1
grep -E "bx_|index_max" gpurun_out/ball_launches_$TAG.csv | tail -6 | awk -F'","' '{print $5, $NF}' | cut -c1-160
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats \
    --clock-control none --profile-from-start off -f -o /tmp/train_$TAG python tools/ncu_step.py train > gpurun_out/ncu_train_$TAG.log 2>&1
ncu -i /tmp/train_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_train_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_train_${TAG}_raw.csv
