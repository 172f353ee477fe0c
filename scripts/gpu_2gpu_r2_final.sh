#!/bin/bash
# Short 2-GPU check of the final tree: bench.py N=2, config 5 sparse N=2 (adaptive overlap per shard), sharded parity.
TAG=${1:-r2fin2}
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e-host > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
  bench_configs.py --workload C5sparse --steps 20 --warmup 5 > gpurun_out/${TAG}_C5sparse_n2.json 2> gpurun_out/${TAG}_C5sparse_n2.err
python - <<PY
import json
for f, keys in (("gpurun_out/${TAG}_bench_n2.json", ("value", "ms_per_step")), ("gpurun_out/${TAG}_C5sparse_n2.json", ("value", "ms_per_step", "verify"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in keys}, d.get("e2e", {}).get("value"))
    except Exception as ex:
        print(f, "failed:", ex, open(f.replace(".json", ".err")).read()[-800:])
PY
NCCL_DEBUG=WARN timeout 150 python -m pytest tests/test_gpu_parity.py::test_random_mixed_batches tests/test_gpu_shards.py -m gpu -q --timeout=120 -k "shards" -x 2>&1 | tail -3
