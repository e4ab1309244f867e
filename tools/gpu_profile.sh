#!/bin/bash
# Run on the GPU box (via gpurun): tests, bench, ncu launch list, one full ncu capture of the dominant kernel.
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 600 python -m pytest tests -m gpu -q --timeout=200 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 --train > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nbatches 2 > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:layer_fwd_tc_kernel -s 40 -c 4 -f -o gpurun_out/prof_tc_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nbatches 2 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -5 gpurun_out/pytest_$TAG.log
tail -c 1800 gpurun_out/bench_$TAG.json
tail -3 gpurun_out/bench_$TAG.err
ls -la gpurun_out | tail -8
