#!/bin/bash
TAG=r2l
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4), "frac", round(d.get("frac_of_hbm_peak", 0), 3),
          d.get("roofline", {}).get("stage_ms"), (d.get("config") or {}).get("verify", d.get("verify")), (d.get("config") or {}).get("output"))
except Exception as ex:
    print(f, "failed:", ex, open(f.replace(".json", ".err")).read()[-500:])
PY
}
timeout 900 python bench.py --pool --no-secondary --no-cpu > gpurun_out/${TAG}_bench_pool.json 2> gpurun_out/${TAG}_bench_pool.err; show gpurun_out/${TAG}_bench_pool.json
timeout 600 python bench_configs.py --workload C3 --steps 10 > gpurun_out/${TAG}_C3_rings_m128.json 2> gpurun_out/${TAG}_C3_rings_m128.err; show gpurun_out/${TAG}_C3_rings_m128.json
timeout 600 python bench_configs.py --workload C3 --steps 10 --pool 30 > gpurun_out/${TAG}_C3_pool_m128.json 2> gpurun_out/${TAG}_C3_pool_m128.err; show gpurun_out/${TAG}_C3_pool_m128.json
timeout 900 python bench_configs.py --workload C3 --steps 6 --msgs 512 --pool 110 > gpurun_out/${TAG}_C3_pool_m512.json 2> gpurun_out/${TAG}_C3_pool_m512.err; show gpurun_out/${TAG}_C3_pool_m512.json
timeout 900 python bench_configs.py --workload C3 --steps 5 --msgs 1024 --pool 125 > gpurun_out/${TAG}_C3_pool_m1024.json 2> gpurun_out/${TAG}_C3_pool_m1024.err; show gpurun_out/${TAG}_C3_pool_m1024.json
