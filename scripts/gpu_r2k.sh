#!/bin/bash
TAG=r2k
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${TAG}_pytest.log 2>&1; tail -12 gpurun_out/${TAG}_pytest.log
timeout 420 python bench_configs.py --workload C4 --steps 10 --warmup 3 > gpurun_out/${TAG}_C4.json 2> gpurun_out/${TAG}_C4.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_C4.json").read().strip().splitlines()[-1])
    print("C4", round(d["msgs_per_s"] / 1e9, 3), "G msgs/s", round(d["ms_per_step"], 4), "ms frac", round(d["frac_of_hbm_peak"], 3), d["roofline"]["stage_ms"])
except Exception as ex:
    print("C4 failed", ex)
PY
timeout 900 python bench.py --no-secondary > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
echo "ref exit $?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "steps")}, d.get("roofline", {}).get("stage_ms"))
        for k in ("sustained", "e2e", "e2e_host", "cpu_baseline"):
            if d.get(k) is not None:
                print("  ", k, json.dumps(d[k])[:600])
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
