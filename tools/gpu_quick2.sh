#!/bin/bash
# new-feature tests first (verbose failures), then the whole GPU suite
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_runtime.py tests/test_gpu_random.py -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_new.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_new.log; tail -15 gpurun_out/pytest_gpu.log
