#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list of one eager step, layer timers.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/lp_time.py > gpurun_out/lp_time.log 2>&1
timeout 300 python tools/ap_time.py > gpurun_out/ap_time.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/lp_time.log gpurun_out/ap_time.log; head -12 gpurun_out/launches.txt
