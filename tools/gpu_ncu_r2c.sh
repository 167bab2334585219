#!/bin/bash
# ncu --set full of the two-product attention kernels at the head of round 2 (K-split forward).  The report stays on the GPU
# box; the raw metric table and its summary come home.
set -u
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none -k regex:"persist4d" -c 8 -f -o /tmp/r02c python tools/ap_time.py \
  > gpurun_out/r02c_ncu.log 2>&1
tail -2 gpurun_out/r02c_ncu.log
ncu -i /tmp/r02c.ncu-rep --page raw --csv > gpurun_out/r02c_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02c_raw.csv > gpurun_out/r02c_ncu_full_persist4d.csv 2>&1
wc -c gpurun_out/r02c_raw.csv
