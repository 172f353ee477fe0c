#!/bin/bash
# multi-GPU ingest A/B: NCCL broadcast per step vs peer-memory ingest (CUDA IPC + TMA loads over NVLink)
N=${1:-2}; TAG=${2:-p2p}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
for mode in p2p nccl; do
  timeout 600 $TR bench_configs.py --workload C5sparse --steps 20 --warmup 3 --mgpu-ingest $mode > gpurun_out/${TAG}_c5sparse_${mode}_n$N.json 2> gpurun_out/${TAG}_c5sparse_${mode}_n$N.err
  grep '^{' gpurun_out/${TAG}_c5sparse_${mode}_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('C5sparse $mode', d['value'], d['ms_per_step'], d['config'].get('parallelism'))" || tail -5 gpurun_out/${TAG}_c5sparse_${mode}_n$N.err
  timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --ingest $mode > gpurun_out/${TAG}_bench_${mode}_n$N.json 2> gpurun_out/${TAG}_bench_${mode}_n$N.err
  grep '^{' gpurun_out/${TAG}_bench_${mode}_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('C2 $mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['verify'], d['config']['parallelism'])" || tail -5 gpurun_out/${TAG}_bench_${mode}_n$N.err
done
