#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02d}
echo "== A: full-size ball test, plain"
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "ball_group_grid_full_size" -p no:cacheprovider 2>&1 | tail -3
echo "== B: same, USIP_BALL_NO_PDL=1"
USIP_BALL_NO_PDL=1 timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "ball_group_grid_full_size" -p no:cacheprovider 2>&1 | tail -3
echo "== C: same under compute-sanitizer memcheck"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "ball_group_grid_full_size" -p no:cacheprovider > gpurun_out/sanitizer_ballfull_$TAG.log 2>&1
grep -m 30 -E "Invalid|at 0x|by thread|Address|ERROR SUMMARY|passed|failed|bx_|Error" gpurun_out/sanitizer_ballfull_$TAG.log | head -30
echo "== D: ablation ball, plain / no PDL"
timeout 600 python -m pytest tests/test_gpu_detector.py -q -x -k "ablation and ball" -p no:cacheprovider 2>&1 | tail -3
USIP_BALL_NO_PDL=1 timeout 600 python -m pytest tests/test_gpu_detector.py -q -x -k "ablation and ball" -p no:cacheprovider 2>&1 | tail -3
echo "== E: racecheck on the ablation ball test"
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_detector.py -q -x -k "ablation and ball and False" -p no:cacheprovider > gpurun_out/racecheck_ball_$TAG.log 2>&1
grep -m 30 -E "Race|hazard|at 0x|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|bx_" gpurun_out/racecheck_ball_$TAG.log | head -30
echo "== F: per-file suites with PDL off (everything else), then bench"
export USIP_BALL_NO_PDL=1
for f in test_gpu_ops test_gpu_detector test_gpu_vs_reference; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_$TAG.err
