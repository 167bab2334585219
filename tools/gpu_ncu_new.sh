#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'bn_apply|dropout_kernel|outer|persist4' -c 10 \
  -f -o gpurun_out/r01_new_kernels python tools/ncu_new_kernels.py > gpurun_out/ncu_new.log 2>&1
tail -3 gpurun_out/ncu_new.log
ncu -i gpurun_out/r01_new_kernels.ncu-rep --page raw --csv > gpurun_out/r01_new_kernels_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r01_new_kernels_raw.csv > gpurun_out/r01_new_kernels_summary.csv 2>&1
cut -c1-200 gpurun_out/r01_new_kernels_summary.csv | head -30
