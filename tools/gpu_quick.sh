#!/bin/bash
# Quick GPU visit: targeted tests first (own timeouts so a hung kernel cannot hold the box), then timers + bench.
set -u
mkdir -p gpurun_out
( time timeout -s KILL 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention_rnn or lstm_layer" ) > gpurun_out/pytest_attn.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_attn.log
tail -15 gpurun_out/pytest_attn.log
timeout -s KILL 200 python tools/lp_time.py > gpurun_out/lp_time.log 2>&1; tail -4 gpurun_out/lp_time.log
timeout -s KILL 200 python tools/ap_time.py > gpurun_out/ap_time.log 2>&1; cat gpurun_out/ap_time.log
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json | cut -c1-300; tail -3 gpurun_out/bench.err
