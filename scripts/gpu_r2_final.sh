#!/bin/bash
# Last pass of round 2: A/B of the adaptive pack-stream overlap (variant bit 5 = off) on the broadcast configs, then the
# full single-GPU validation (gpu_full.sh) on the same box.
TAG=${1:-r2fin}
mkdir -p gpurun_out
for wl in C5sparse C5dense C3; do
  for v in 0 32; do
    timeout 300 python bench_configs.py --workload $wl --steps 12 --warmup 4 --variant $v > gpurun_out/${TAG}_cfg_${wl}_v$v.json 2> gpurun_out/${TAG}_cfg_${wl}_v$v.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_cfg_${wl}_v$v.json").read().strip().splitlines()[-1])
    print("$wl v$v", round(d["value"], 1), "GB/s", round(d["ms_per_step"], 4), "ms", "frac", round(d["frac_of_hbm_peak"], 3), d.get("verify"))
except Exception as ex:
    print("$wl v$v failed", ex, open("gpurun_out/${TAG}_cfg_${wl}_v$v.err").read()[-600:])
PY
  done
done
bash scripts/gpu_full.sh $TAG
