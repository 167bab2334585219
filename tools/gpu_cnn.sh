#!/bin/bash
# CNN front-end: parity tests, then its forward + backward time at three batch sizes
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_cnn.py tests/test_gpu_runtime.py -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_new.log 2>&1
timeout 300 python tools/cnn_time.py > gpurun_out/cnn_time.log 2>&1
tail -12 gpurun_out/pytest_new.log; tail -5 gpurun_out/cnn_time.log
