#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 800 python -m pytest tests/test_gpu_cnn.py tests/test_gpu_runtime.py -q -m gpu 2>&1 | tail -60 ) > gpurun_out/pytest_new.log 2>&1
cat gpurun_out/pytest_new.log
