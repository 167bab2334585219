#!/bin/bash
# Round 2 verification: the whole -m gpu suite, the bench line of every BASELINE configuration, the launch list of
# the reference-default graph (ncu durations are serialised / cold-cache: only shares compare with the graph-timed step).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu.log
for c in 1 2 3 4; do
  timeout 300 python bench.py --config $c --skip-cpu-baseline --skip-extras > gpurun_out/r2_bench_cfg$c.json 2> gpurun_out/r2_bench_cfg$c.err
  cut -c1-400 gpurun_out/r2_bench_cfg$c.json
done
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
cut -c1-300 gpurun_out/r2_bench_default.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02_launches.csv python tools/profile_step.py --graph default > gpurun_out/r02_profile_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_default_graph.txt 2>&1
head -40 gpurun_out/r02_launches_default_graph.txt
