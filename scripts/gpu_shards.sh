#!/bin/bash
# Runs on a box with N >= 2 GPUs (gpurun --gpus N): the sharded-engine parity tests with the library's
# NCCL ingest, then bench.py under torchrun through the same entry point.
# usage: scripts/gpu_shards.sh [tag] [ngpus]
TAG=${1:-r2}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_shards.py "tests/test_gpu_parity.py::test_random_mixed_batches" -m gpu -q --timeout=600 -k "shards" > gpurun_out/${TAG}_pytest_shards.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_shards.log
tail -5 gpurun_out/${TAG}_pytest_shards.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  fi
  echo "bench n=$n exit $?"
  tail -3 gpurun_out/${TAG}_bench_n$n.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n$n.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")}, d["e2e"]["value"], d["config"]["verify"], d["config"]["parallelism"])
except Exception as ex:
    print("no bench line:", ex)
PY
done
