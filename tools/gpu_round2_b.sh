#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_capture.py > gpurun_out/debug_capture.log 2>&1
tail -40 gpurun_out/debug_capture.log
USIP_NO_TRAIN_GRAPH=1 timeout 1200 python -m pytest tests/test_gpu_vs_reference.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/pytest_vsref.log 2>&1
tail -5 gpurun_out/pytest_vsref.log
USIP_NO_TRAIN_GRAPH=1 timeout 1200 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider --deselect tests/test_gpu_vs_reference.py > gpurun_out/pytest_nograph.log 2>&1
tail -5 gpurun_out/pytest_nograph.log
