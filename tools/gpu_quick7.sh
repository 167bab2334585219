#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_cnn.py -q -m gpu -k "alone" 2>&1 | tail -15 ) > gpurun_out/pytest_new.log 2>&1
timeout 300 python tools/cnn_time.py > gpurun_out/cnn_time.log 2>&1
cat gpurun_out/pytest_new.log | tail -8; cat gpurun_out/cnn_time.log | tail -6
