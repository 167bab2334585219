#!/bin/bash
# ncu launch list of one eager training step of the bench workload (durations are serialised and cold-cache)
set -u
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1
cat gpurun_out/launches.txt | head -40
