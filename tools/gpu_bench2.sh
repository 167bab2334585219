#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_runtime.py -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_new.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
cat gpurun_out/pytest_new.log | tail -5; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','reference_default_graph','e2e_tfrecord','cpu_baseline','clocks'):
    print(k, l.get(k))
print('roofline', {k:l['roofline'].get(k) for k in ('achieved','peak','frac')} if l.get('roofline') else None)
PY
