#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cnn.py tests/test_gpu_runtime.py -q -m gpu -x > gpurun_out/r2_cnn_tests.log 2>&1
tail -5 gpurun_out/r2_cnn_tests.log
timeout 300 python tools/cnn_time.py > gpurun_out/r2_cnn_time.log 2>&1
tail -3 gpurun_out/r2_cnn_time.log
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/r2_bench_default2.json 2> gpurun_out/r2_bench_default2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d.get('with_resnet_cnn')))
print({k:(v.get('value'),v.get('ms_per_step'),v.get('gpu_launches_per_step')) for k,v in d['configs'].items()})
PY
tail -3 gpurun_out/r2_bench_default2.err
