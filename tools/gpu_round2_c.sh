#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_detector.py -q -x -k "ablation and ball and False" -p no:cacheprovider > gpurun_out/sanitizer_ball_$TAG.log 2>&1
grep -m 40 -E "Invalid|at 0x|by thread|Address|ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_ball_$TAG.log | head -40
timeout 600 python tools/debug_capture.py > gpurun_out/debug_capture.log 2>&1
grep -E "OK|FAILED" gpurun_out/debug_capture.log
for f in test_gpu_ops test_gpu_detector test_gpu_vs_reference test_dropin_imports test_gpu_dp_nccl; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
tail -c 2500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_$TAG.err
