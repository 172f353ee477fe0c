#!/bin/bash
TAG=r2n
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -k "pool or runs or configs" > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --pool --no-secondary --no-cpu > gpurun_out/${TAG}_bench_pool.json 2> gpurun_out/${TAG}_bench_pool.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_pool.json").read().strip().splitlines()[-1])
    print("C2 pool", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), d["roofline"]["stage_ms"], "pack frac", round(d["roofline"]["frac"], 4), d["config"]["verify"], "e2e", round(d["e2e"]["value"], 1), "e2e_host", round(d["e2e_host"]["value"], 1))
except Exception as ex:
    print("pool bench failed", ex, open("gpurun_out/${TAG}_bench_pool.err").read()[-800:])
PY
bash scripts/gpu_r2m.sh
