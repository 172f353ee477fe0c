#!/bin/bash
# A/B of the adaptive pack-stream overlap (variant bit 5 = off) on the broadcast configs + the C2 bench line.
TAG=${1:-r2ov}
mkdir -p gpurun_out
for wl in C5sparse C5dense; do
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
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --timeout=300 2>&1 | tail -2
timeout 600 python bench.py --no-secondary --no-e2e-host > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d["sustained"]["value"], d["e2e"]["value"], d["roofline"]["frac"])
PY
