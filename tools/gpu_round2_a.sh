#!/bin/bash
# round-2 GPU pass A: full -m gpu suite (incl. parity against the reference itself), default bench line, the reference arm,
# and ONE ncu --set full pass over every kernel of the step (raw CSV comes back, the .ncu-rep stays if it is small).
mkdir -p gpurun_out
TAG=${1:-r02a}
timeout 1500 python -m pytest tests -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1
tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
tail -c 1500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_$TAG.err
# (1) --set full on the descriptor path + stand-alone operators (bg_*, layer kernels of the descriptor, index_max, ball_query)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/descops_$TAG python tools/ncu_step.py desc ops > gpurun_out/ncu_descops_$TAG.log 2>&1
tail -2 gpurun_out/ncu_descops_$TAG.log
ncu -i /tmp/descops_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_descops_${TAG}_raw.csv 2>/dev/null
sz=$(stat -c %s /tmp/descops_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -lt 30000000 ] && [ "$sz" -gt 0 ]; then cp /tmp/descops_$TAG.ncu-rep gpurun_out/; fi
# (2) the sections the roofline argument needs on every kernel of one detector train step (fwd + loss + backward + Adam)
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats \
    --clock-control none --profile-from-start off -f -o /tmp/train_$TAG python tools/ncu_step.py train > gpurun_out/ncu_train_$TAG.log 2>&1
tail -2 gpurun_out/ncu_train_$TAG.log
ncu -i /tmp/train_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_train_${TAG}_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out/*_raw.csv
