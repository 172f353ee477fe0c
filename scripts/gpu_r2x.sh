#!/bin/bash
# last pass of the round: the whole GPU suite on the final tree + the small regression / evidence runs
TAG=r2x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 python bench_configs.py --workload C1 > gpurun_out/${TAG}_C1.json 2> gpurun_out/${TAG}_C1.err; cut -c1-600 gpurun_out/${TAG}_C1.json
timeout 300 python bench_configs.py --workload churn --steps 10 > gpurun_out/${TAG}_churn.json 2> gpurun_out/${TAG}_churn.err; cut -c1-500 gpurun_out/${TAG}_churn.json
timeout 300 python bench.py --conns 65536 --host-rings --no-cpu --no-secondary --steps 10 > gpurun_out/${TAG}_hostrings.json 2> gpurun_out/${TAG}_hostrings.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_hostrings.json").read().strip().splitlines()[-1])
    print("host rings 64K conns: value", round(d["value"], 1), "GB/s e2e", round(d["e2e"]["value"], 1), "e2e_host", round(d["e2e_host"]["value"], 1), d["config"]["verify"])
except Exception as ex:
    print("host rings failed", ex, open("gpurun_out/${TAG}_hostrings.err").read()[-600:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -f -k "regex:k_pool_finish|k_offsets" -s 8 -c 2 -o gpurun_out/${TAG}_pool python bench.py --pool --steps 2 --warmup 3 --no-cpu --no-verify --no-secondary --no-e2e-host --sustain 0 > gpurun_out/${TAG}_pool.log 2>&1
ls -la gpurun_out | grep ${TAG}_pool
