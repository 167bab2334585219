#!/bin/bash
# ncu --set full of the kernels added in round 2 (dual-attention decoder, tensor-core convolutions).  The report stays on
# the GPU box (it exceeds the 64 MiB that travel back); the raw metric table and its summary come home.
set -u
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --profile-from-start off -k regex:"wlas|conv_mma" -c 40 \
  -f -o /tmp/r02_new_kernels python tools/ncu_r2_new_kernels.py > gpurun_out/r02_ncu_new.log 2>&1
tail -2 gpurun_out/r02_ncu_new.log
ncu -i /tmp/r02_new_kernels.ncu-rep --page raw --csv > gpurun_out/r02_new_kernels_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_new_kernels_raw.csv > gpurun_out/r02_ncu_full_wlas_conv.csv 2>&1
ls -la gpurun_out | head
