#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_smoke.py > gpurun_out/dp_smoke.log 2>&1; echo "dp_smoke exit $?" >> gpurun_out/dp_smoke.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench2 exit $?" >> gpurun_out/bench2.err
grep -v "^\[rank 1" gpurun_out/dp_smoke.log | tail -14; tail -2 gpurun_out/bench2.err; cut -c1-700 gpurun_out/bench2.json
