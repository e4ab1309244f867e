#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02g}
timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "ball or index_max or wgrad" -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_detector.py -q -k "ablation or descriptor" -p no:cacheprovider 2>&1 | tail -3
python tools/bench_descriptor.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('ball_group_fused','ball_group_reference_gpu','index_max_op','descriptor_forward_eval','descriptor_train_step') if k in d})"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bx_|index_max" -f -o /tmp/ball_$TAG python tools/ncu_step.py ballonly > gpurun_out/ncu_ball_$TAG.log 2>&1
ncu -i /tmp/ball_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_ball_${TAG}_raw.csv 2>/dev/null
sz=$(stat -c %s /tmp/ball_$TAG.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -lt 30000000 ] && [ "$sz" -gt 0 ]; then cp /tmp/ball_$TAG.ncu-rep gpurun_out/; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ball_launches_$TAG.csv python tools/ncu_step.py ballonly > /dev/null 2>&1
grep -E "bx_|index_max" gpurun_out/ball_launches_$TAG.csv | tail -6 | awk -F'","' '{print substr($5,1,40), $NF}'
