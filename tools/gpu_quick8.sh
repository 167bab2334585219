#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/cnn_time.py > gpurun_out/cnn_time.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; tail -6 gpurun_out/cnn_time.log
