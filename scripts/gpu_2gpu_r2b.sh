#!/bin/bash
TAG=${1:-r2r}
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_shards.py tests/test_c_abi.py "tests/test_gpu_parity.py::test_random_mixed_batches" -m gpu -q --timeout=600 -k "shards or c_host" > gpurun_out/${TAG}_pytest_shards.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_shards.log
tail -4 gpurun_out/${TAG}_pytest_shards.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e-host > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
python - <<PY
import json
f = "gpurun_out/${TAG}_bench_n2.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("bench n2 value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["config"]["verify"])
except Exception as ex:
    print(f, "failed:", ex, open(f.replace(".json", ".err")).read()[-800:])
PY
for n in 2 1; do
  if [ $n -eq 1 ]; then CMD="python"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532"; fi
  timeout 900 $CMD bench_configs.py --workload C5sparse --steps 20 --warmup 5 > gpurun_out/${TAG}_C5sparse_n$n.json 2> gpurun_out/${TAG}_C5sparse_n$n.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_C5sparse_n$n.json").read().strip().splitlines()[-1])
    print("C5sparse n=$n", round(d["value"], 1), "GB/s", round(d["ms_per_step"], 4), "ms/step")
except Exception as ex:
    print("C5sparse n=$n failed", ex, open("gpurun_out/${TAG}_C5sparse_n$n.err").read()[-800:])
PY
done
