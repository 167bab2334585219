#!/bin/bash
# the driver's default bench run, timed end to end (wall clock), with a short digest of the line
set -u
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2_bench_default3.json 2> gpurun_out/r2_bench_default3.err
echo "bench.py wall seconds: $(( $(date +%s) - t0 ))"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_default3.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"])
print("tfrecord", json.dumps(d.get("e2e_tfrecord"))[:300])
print("cnn", json.dumps(d.get("with_resnet_cnn"))[:400])
print("cpu", json.dumps(d.get("cpu_baseline"))[:300])
print({k: (v.get("value"), v.get("ms_per_step"), (v.get("roofline") or {}).get("frac")) for k, v in d["configs"].items()})
print("gate", d["roofline_tensor"]["gate_gemms"]["frac"], d["roofline_tensor"]["gate_gemms"].get("frac_of_roofline"))
PY
tail -3 gpurun_out/r2_bench_default3.err
