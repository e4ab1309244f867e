#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02e}
echo "== layer kernel variants"
timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "layer_fwd or wgrad" -p no:cacheprovider 2>&1 | tail -4
for f in test_gpu_ops test_gpu_detector test_gpu_vs_reference test_dropin_imports; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
echo "== bench (A through TMEM, default)"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r02e.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('train_step'), j['roofline'].get('per_op_ms'))
d=j.get('descriptor',{}); print({k:d.get(k) for k in ('ball_group_fused','index_max_op','descriptor_forward_eval')})
PY
tail -3 gpurun_out/bench_$TAG.err
echo "== bench (A through shared memory, round-1 variant)"
USIP_TC_SMEM_A=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-descriptor > gpurun_out/bench_${TAG}_smemA.json 2> gpurun_out/bench_${TAG}_smemA.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r02e_smemA.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('train_step'), j['roofline'].get('per_op_ms'))
PY
echo "== ball group launch list (PDL on / off)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ball_launches_$TAG.csv python tools/ncu_step.py ballonly > /dev/null 2>&1
USIP_BALL_NO_PDL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ball_launches_nopdl_$TAG.csv python tools/ncu_step.py ballonly > /dev/null 2>&1
grep -E "bx_|index_max" gpurun_out/ball_launches_$TAG.csv | tail -6 | cut -c1-200
USIP_BALL_NO_PDL=1 python tools/bench_descriptor.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-PDL', d['ball_group_fused']['ms_median'], d['index_max_op'])"
