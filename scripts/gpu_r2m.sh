#!/bin/bash
TAG=r2m
mkdir -p gpurun_out
for v in 0 65536 131072 196608; do   # tiles per grab: 1, 2, 4, 8
  timeout 420 python bench_configs.py --workload C5sparse --steps 20 --warmup 3 --variant $v > gpurun_out/${TAG}_C5s_v$v.json 2> gpurun_out/${TAG}_C5s_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_C5s_v$v.json").read().strip().splitlines()[-1])
    print("C5sparse grab variant $v", round(d["value"], 1), "GB/s", round(d["ms_per_step"], 4), "ms", d["roofline"]["stage_ms"], "pack frac", round(d["roofline"]["frac"], 3))
except Exception as ex:
    print("failed $v", ex)
PY
done
