#!/bin/bash
# dual-attention persistent kernels: op-level parity, model-level config 4 cases, config 4 bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "dropout_persistent or attention_rnn_fwd_bwd" > gpurun_out/r2_wlas_ops.log 2>&1
tail -15 gpurun_out/r2_wlas_ops.log
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_random.py -m gpu -q -k "4 or default" > gpurun_out/r2_wlas_model.log 2>&1
tail -15 gpurun_out/r2_wlas_model.log
timeout 300 python bench.py --config 4 --skip-cpu-baseline --skip-extras > gpurun_out/r2_bench_cfg4b.json 2> gpurun_out/r2_bench_cfg4b.err
cut -c1-300 gpurun_out/r2_bench_cfg4b.json; tail -3 gpurun_out/r2_bench_cfg4b.err
