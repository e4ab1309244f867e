#!/bin/bash
# backward pass of the round-2 loop: wgrad micro-benchmark (A/B against usip_b200/lib/libusip_b200_prev.so when that older build
# is present), nearest-neighbour micro-benchmark, GPU tests, default bench, ball-group launch list, ncu capture of one train step
mkdir -p gpurun_out
TAG=${1:-r02j}
python tools/wgrad_microbench.py > gpurun_out/wgrad_new_$TAG.json 2> gpurun_out/wgrad_new_$TAG.err
[ -f usip_b200/lib/libusip_b200_prev.so ] && python tools/wgrad_microbench.py usip_b200/lib/libusip_b200_prev.so > gpurun_out/wgrad_prev_$TAG.json 2> gpurun_out/wgrad_prev_$TAG.err
python - <<PY
import json
try:
    a=json.load(open('gpurun_out/wgrad_new_$TAG.json')); b=json.load(open('gpurun_out/wgrad_prev_$TAG.json'))
    for k in a: print(k, 'new', a[k]['p1_us'], a[k]['p4_us'], 'prev', b[k]['p1_us'], b[k]['p4_us'], 'err', a[k]['p1_err'], a[k]['p4_err'], 'floor', a[k]['mma_floor_us_at_1965MHz'])
except Exception as e: print('ERR', e)
PY
tail -2 gpurun_out/wgrad_new_$TAG.err
python tools/nn_microbench.py > gpurun_out/nn_micro_$TAG.json 2> gpurun_out/nn_micro_$TAG.err; cat gpurun_out/nn_micro_$TAG.json; tail -2 gpurun_out/nn_micro_$TAG.err
for f in test_gpu_ops test_gpu_detector test_gpu_vs_reference; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
grep -h "AssertionError\|Error" gpurun_out/pytest_test_gpu_*_$TAG.log | head -10 | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step')}); print(j.get('train_step')); print(j.get('train_step_tf32_backward'))
d=j.get('descriptor',{}); print({k:d.get(k) for k in ('ball_group_fused','index_max_op','descriptor_forward_eval','descriptor_train_step')})
print(j.get('reference_gpu',{}).get('tf32_off'))
PY
tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ball_launches_$TAG.csv python tools/ncu_step.py ballonly > /dev/null 2>&1
grep -E "bx_|index_max" gpurun_out/ball_launches_$TAG.csv | tail -6 | awk -F'","' '{print substr($5,1,40), $NF}'
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats \
    --clock-control none --profile-from-start off -f -o /tmp/train_$TAG python tools/ncu_step.py train > gpurun_out/ncu_train_$TAG.log 2>&1
ncu -i /tmp/train_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_train_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_train_${TAG}_raw.csv
