#!/bin/bash
# multi-GPU: config 5 (sparse + dense shards) and the C2 bench under torchrun on N GPUs of one box
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench_configs.py --workload C5sparse --steps 20 --warmup 3 > gpurun_out/${TAG}_c5sparse_n$N.json 2> gpurun_out/${TAG}_c5sparse_n$N.err
grep '^{' gpurun_out/${TAG}_c5sparse_n$N.json | cut -c 1-700; tail -3 gpurun_out/${TAG}_c5sparse_n$N.err
timeout 600 $TR bench_configs.py --workload C5dense --steps 10 --warmup 3 > gpurun_out/${TAG}_c5dense_n$N.json 2> gpurun_out/${TAG}_c5dense_n$N.err
grep '^{' gpurun_out/${TAG}_c5dense_n$N.json | cut -c 1-700; tail -3 gpurun_out/${TAG}_c5dense_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
grep '^{' gpurun_out/${TAG}_bench_n$N.json | cut -c 1-900; tail -3 gpurun_out/${TAG}_bench_n$N.err
