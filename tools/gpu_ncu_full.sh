#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:persist4 -c 4 \
  -f -o gpurun_out/r01_persist4 python tools/ncu_layers.py > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
