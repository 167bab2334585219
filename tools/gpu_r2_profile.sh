#!/bin/bash
# Round 2 profiles of the reference-default graph: (1) ncu launch list of one eager training step (durations are
# serialised and cold-cache: only the SHARES are comparable with the graph-timed step), (2) ncu --set full of the
# persistent kernels under dropout (attn_persist4d.cu two-product kernels, lstm_persist4 DROP variants).
set -u
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02_launches.csv python tools/profile_step.py --graph default > gpurun_out/r02_profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_default_graph.txt 2>&1
head -30 gpurun_out/r02_launches_default_graph.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:persist4 -c 4 \
  -f -o gpurun_out/r02_persist4d python tools/ncu_layers.py --drop > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_ncu_full.log
ncu -i gpurun_out/r02_persist4d.ncu-rep --page raw --csv > gpurun_out/r02_persist4d_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_persist4d_raw.csv > gpurun_out/r02_ncu_full_persist4d.csv 2>&1
cut -c1-260 gpurun_out/r02_ncu_full_persist4d.csv
