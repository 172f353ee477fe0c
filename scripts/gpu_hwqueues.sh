#!/bin/bash
# Does head-of-line blocking in the hardware work queues (streams aliased onto CUDA_DEVICE_MAX_CONNECTIONS = 8 channels) explain
# why the pack-stream overlap works when the host enqueues late (scripts/gpu_timeline.sh) and not when everything is queued ahead?
TAG=${1:-hwq}
mkdir -p gpurun_out
run() {  # name, connections, workload, variant
  CUDA_DEVICE_MAX_CONNECTIONS=$2 timeout 60 python bench_configs.py --workload $3 --steps 12 --warmup 4 --variant $4 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1: connections=$2 $3 variant $4 ->", round(d["value"], 1), "GB/s", round(d["ms_per_step"], 4), "ms", d.get("verify", "")[:40])
except Exception as ex:
    print("$1 failed", ex, open("gpurun_out/${TAG}_$1.err").read()[-400:])
PY
}
run sparse_c32 32 C5sparse 0
run sparse_c8 8 C5sparse 0
run dense_forced_c32 32 C5dense 8
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 90 python bench.py --no-secondary --no-e2e-host > gpurun_out/${TAG}_bench_c32.json 2> gpurun_out/${TAG}_bench_c32.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_c32.json").read().strip().splitlines()[-1])
print("C2 bench with 32 connections:", {k: d.get(k) for k in ("value", "ms_per_step")}, d["sustained"]["value"], d["e2e"]["value"], d["roofline"]["frac"])
PY
