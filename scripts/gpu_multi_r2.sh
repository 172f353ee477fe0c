#!/bin/bash
# N-GPU box (gpurun --gpus N): sharded parity tests with the library's NCCL ingest, bench.py under torchrun at 2..N,
# config 5 (sparse and dense) through the sharded engine at N.
TAG=${1:-r2}
N=${2:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_shards.py tests/test_c_abi.py "tests/test_gpu_parity.py::test_random_mixed_batches" -m gpu -q --timeout=600 -k "shards or c_host" > gpurun_out/${TAG}_pytest_shards.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_shards.log
tail -4 gpurun_out/${TAG}_pytest_shards.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], "value", round(d["value"], 1), "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4),
          "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None, "e2e_host", round(d["e2e_host"]["value"], 1) if d.get("e2e_host") else None,
          (d.get("config") or {}).get("verify"), (d.get("config") or {}).get("parallelism", "")[:90])
except Exception as ex:
    print(sys.argv[1], "failed:", ex, open(sys.argv[1].replace(".json", ".err")).read()[-600:])
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; show gpurun_out/${TAG}_bench_n1.json
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
    bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  show gpurun_out/${TAG}_bench_n$n.json
done
for wl in C5sparse C5dense; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench_configs.py --workload $wl --steps 20 --warmup 5 > gpurun_out/${TAG}_${wl}_n$N.json 2> gpurun_out/${TAG}_${wl}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${wl}_n$N.json").read().strip().splitlines()[-1])
    print("$wl n=$N", round(d["value"], 1), "GB/s", round(d["ms_per_step"], 4), "ms/step", d["config"].get("parallelism", "")[:100])
except Exception as ex:
    print("$wl failed", ex, open("gpurun_out/${TAG}_${wl}_n$N.err").read()[-600:])
PY
done
timeout 600 python bench_configs.py --workload C5sparse --steps 20 --warmup 5 > gpurun_out/${TAG}_C5sparse_n1.json 2> gpurun_out/${TAG}_C5sparse_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_C5sparse_n1.json").read().strip().splitlines()[-1])
print("C5sparse n=1 (same box)", round(d["value"], 1), "GB/s", round(d["ms_per_step"], 4), "ms/step")
PY
