#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cnn.py -q -m gpu -x > gpurun_out/r2_cnn_tests.log 2>&1
tail -25 gpurun_out/r2_cnn_tests.log
timeout 300 python tools/cnn_time.py > gpurun_out/r2_cnn_time.log 2>&1
tail -4 gpurun_out/r2_cnn_time.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_cnn_launches2.csv python tools/cnn_profile.py 256 > gpurun_out/r02_cnn_profile.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_cnn_launches2.csv | head -30
