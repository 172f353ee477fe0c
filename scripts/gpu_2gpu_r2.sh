#!/bin/bash
# 2-GPU box: sharded parity (library NCCL ingest), C host on two GPUs, bench N=2 (rings and pool)
TAG=${1:-r2p}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_shards.py tests/test_c_abi.py tests/test_gpu_hook_sync.py "tests/test_gpu_parity.py::test_random_mixed_batches" -m gpu -q --timeout=600 -k "shards or c_host" > gpurun_out/${TAG}_pytest_shards.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_shards.log
tail -4 gpurun_out/${TAG}_pytest_shards.log
for extra in "" "--pool"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 5 $extra > gpurun_out/${TAG}_bench_n2${extra}.json 2> gpurun_out/${TAG}_bench_n2${extra}.err
  python - <<PY
import json
f = "gpurun_out/${TAG}_bench_n2${extra}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "e2e_host", round(d["e2e_host"]["value"], 1), d["config"]["verify"], d["config"]["parallelism"][:80])
except Exception as ex:
    print(f, "failed:", ex, open(f.replace(".json", ".err")).read()[-800:])
PY
done
