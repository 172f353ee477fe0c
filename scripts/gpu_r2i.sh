#!/bin/bash
TAG=r2i
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench_configs.py --workload latency > gpurun_out/${TAG}_latency.json 2> gpurun_out/${TAG}_latency.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_latency.json").read().strip().splitlines()[-1])
    for c in d["cases"]:
        print(round(c["p50_us"], 1), "us p50 ", c["case"][:110])
except Exception as ex:
    print("latency failed", ex)
PY
# direct path: pack-stream overlap x CTAs per SM of the direct pack
for v in 0 8 24584 20488 16392; do   # 0 | 8 | 8+(6<<12) | 8+(5<<12) | 8+(4<<12)
  timeout 420 python bench_configs.py --workload C4 --steps 10 --warmup 3 --variant $v > gpurun_out/${TAG}_C4_v$v.json 2> gpurun_out/${TAG}_C4_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_C4_v$v.json").read().strip().splitlines()[-1])
    print("C4 variant $v", round(d["msgs_per_s"] / 1e9, 3), "G msgs/s", round(d["ms_per_step"], 4), "ms", d["roofline"]["stage_ms"])
except Exception as ex:
    print("C4 v$v failed", ex)
PY
done
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
echo "ref exit $?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "steps")})
        for k in ("sustained", "e2e", "e2e_host", "cpu_baseline", "secondary"):
            if d.get(k) is not None:
                print("  ", k, json.dumps(d[k])[:900])
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
