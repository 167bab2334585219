#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_edge.py -q 2>&1 | tail -40 ) > gpurun_out/pytest_new.log 2>&1
timeout 120 python tools/h2d_bw.py > gpurun_out/h2d_bw.log 2>&1
cat gpurun_out/pytest_new.log; cat gpurun_out/h2d_bw.log
