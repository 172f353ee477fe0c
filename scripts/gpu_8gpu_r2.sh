#!/bin/bash
# 8-GPU box: bench N=8 and config 5 (sparse, dense) through the sharded engine
TAG=${1:-r2q}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err
python - <<PY
import json
f = "gpurun_out/${TAG}_bench_n8.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("bench n8 value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "e2e_host", round(d["e2e_host"]["value"], 1), d["config"]["verify"], d["config"]["parallelism"][:90], "setup_s", d["config"]["setup_s"])
except Exception as ex:
    print(f, "failed:", ex, open(f.replace(".json", ".err")).read()[-1200:])
PY
for wl in C5sparse C5dense; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench_configs.py --workload $wl --steps 20 --warmup 5 > gpurun_out/${TAG}_${wl}_n8.json 2> gpurun_out/${TAG}_${wl}_n8.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${wl}_n8.json").read().strip().splitlines()[-1])
    print("$wl n=8", round(d["value"], 1), "GB/s", round(d["ms_per_step"], 4), "ms/step", d["config"].get("parallelism", "")[:100])
except Exception as ex:
    print("$wl failed", ex, open("gpurun_out/${TAG}_${wl}_n8.err").read()[-1200:])
PY
done
