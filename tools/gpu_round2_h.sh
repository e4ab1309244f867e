#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02h}
timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "ball or index_max" -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_detector.py -q -k "ablation or descriptor" -p no:cacheprovider 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-250
python tools/bench_descriptor.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('ball_group_fused','index_max_op','descriptor_forward_eval') if k in d})"
USIP_BALL_NO_PDL=1 python tools/bench_descriptor.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-PDL', d['ball_group_fused']['ms_median'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ball_launches_$TAG.csv python tools/ncu_step.py ballonly > /dev/null 2>&1
grep -E "bx_|index_max" gpurun_out/ball_launches_$TAG.csv | tail -8 | awk -F'","' '{print substr($5,1,40), $NF}'
