#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py -q -k "other_optimisers or three_training" 2>&1 | tail -30 ) > gpurun_out/pytest_new.log 2>&1
cat gpurun_out/pytest_new.log
