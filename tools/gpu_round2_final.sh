#!/bin/bash
# final pass of round 2: every -m gpu test file, smoke(), the default bench line, the reference arm, ncu evidence
mkdir -p gpurun_out
TAG=${1:-r02final}
for f in test_gpu_ops test_gpu_detector test_gpu_vs_reference test_dropin_imports; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
grep -h "AssertionError\|Error" gpurun_out/pytest_test_*_$TAG.log | head -10 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
SECONDS=0
timeout 1200 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
echo "bench default rc=$? [${SECONDS}s]"
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step','steps','warmup','gpu_launches')}, 'e2e', j['e2e']['value']); print(j.get('train_step')); print(j.get('train_step_tf32_backward'))
print(j.get('roofline',{}).get('kernel'), j.get('roofline',{}).get('frac'), j.get('clocks'))
print(j.get('cpu_baseline'))
d=j.get('descriptor',{}); print({k:d.get(k) for k in ('ball_group_fused','index_max_op','descriptor_forward_eval','descriptor_train_step')})
print(j.get('reference_gpu'))
PY
tail -3 gpurun_out/bench_$TAG.err
SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_$TAG.err
echo "reference arm rc=$? [${SECONDS}s]"; tail -1 gpurun_out/bench_ref_${TAG}.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-train --no-reference-gpu --no-descriptor > /dev/null 2>&1
ls -la gpurun_out/launches_$TAG.csv
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats \
    --clock-control none --profile-from-start off -f -o /tmp/train_$TAG python tools/ncu_step.py fwd train > gpurun_out/ncu_train_$TAG.log 2>&1
ncu -i /tmp/train_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_fwd_train_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_fwd_train_${TAG}_raw.csv
