#!/bin/bash
# Full single-GPU pass on the box: every -m gpu test, smoke, the default bench (all legs) and the reference arm.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E 'passed|failed|FAILED|Error' gpurun_out/${TAG}_pytest.log | tail -20
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
echo "ref exit $?"; tail -2 gpurun_out/${TAG}_ref.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "steps")})
        for k in ("sustained", "e2e", "e2e_host", "cpu_baseline", "secondary"):
            if d.get(k) is not None:
                print("  ", k, json.dumps(d[k])[:700])
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
