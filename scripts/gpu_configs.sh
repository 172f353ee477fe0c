#!/bin/bash
# Secondary configs (C4 direct, C5 sparse, C3) with A/B of the pack-stream overlap (variant 8).
TAG=${1:-r2}
mkdir -p gpurun_out
for wl in C4 C5sparse C3; do
  for v in 0 8; do
    timeout 420 python bench_configs.py --workload $wl --steps 10 --warmup 3 --variant $v > gpurun_out/${TAG}_cfg_${wl}_v$v.json 2> gpurun_out/${TAG}_cfg_${wl}_v$v.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_cfg_${wl}_v$v.json").read().strip().splitlines()[-1])
    print("$wl v$v", round(d["value"], 1), "GB/s", round(d["msgs_per_s"] / 1e9, 3), "G msgs/s", round(d["ms_per_step"], 4), "ms", "frac", round(d["frac_of_hbm_peak"], 3), d["roofline"]["stage_ms"])
except Exception as ex:
    print("$wl v$v failed", ex, open("gpurun_out/${TAG}_cfg_${wl}_v$v.err").read()[-600:])
PY
  done
done
