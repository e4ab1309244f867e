#!/bin/bash
# Run on the GPU box (via gpurun): tests, bench lines, ncu launch list, one full ncu capture of the dominant kernel
# (the 512->512 kNN-fusion layer on 131072 rows, the kernel bench.py reports in `roofline`), descriptor-path numbers.
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --steps 30 --warmup 5 --train --no-cpu-baseline > gpurun_out/bench_${TAG}_1gpu_train.json 2>> gpurun_out/bench_$TAG.err
timeout 600 python tools/bench_descriptor.py 2>&1 | tail -1 > gpurun_out/bench_${TAG}_descriptor.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nbatches 2 > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:layer_fwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_tc_$TAG \
    python tools/tc_one.py 131072 512 512 16 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log
tail -c 600 gpurun_out/bench_${TAG}_1gpu.json
tail -c 400 gpurun_out/bench_${TAG}_1gpu_train.json
tail -3 gpurun_out/bench_$TAG.err
ls -la gpurun_out | tail -8
