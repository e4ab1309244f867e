#!/bin/bash
# full ncu capture of the four wide kNN-fusion layer launches of one step (pair kernel and single-CTA kernel)
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:layer_fwd_tc -s 32 -c 5 -f -o gpurun_out/prof_big_pair_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nbatches 2 --no-graph > gpurun_out/ncu_big_pair_$TAG.log 2>&1
USIP_TC_SINGLE_CTA=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:layer_fwd_tc -s 32 -c 5 -f -o gpurun_out/prof_big_single_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nbatches 2 --no-graph > gpurun_out/ncu_big_single_$TAG.log 2>&1
tail -3 gpurun_out/ncu_big_pair_$TAG.log gpurun_out/ncu_big_single_$TAG.log
ls -la gpurun_out/*.ncu-rep
