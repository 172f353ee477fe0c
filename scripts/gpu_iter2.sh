#!/bin/bash
TAG=${1:-it2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x --timeout=600 2>&1 | tail -3
timeout 300 python bench_configs.py --workload C4 --hit 0.9 --steps 10 --warmup 3 > gpurun_out/${TAG}_cfg_C4_hit90.json 2>>gpurun_out/${TAG}.err
timeout 300 python bench.py --impl reference --conns 128 --msgs 1 --steps 10000 --warmup 100 > gpurun_out/${TAG}_cfg_C1.json 2>>gpurun_out/${TAG}.err
timeout 300 python bench.py --impl reference --conns 2 --payload 10000 --msgs 1 --steps 10000 --warmup 100 >> gpurun_out/${TAG}_cfg_C1.json 2>>gpurun_out/${TAG}.err
timeout 300 python bench_configs.py --workload latency > gpurun_out/${TAG}_latency.json 2>>gpurun_out/${TAG}.err
timeout 300 python bench_configs.py --workload C5sparse --steps 10 > gpurun_out/${TAG}_c5s_n1.json 2>>gpurun_out/${TAG}.err
head -c 1500 gpurun_out/${TAG}_cfg_C4_hit90.json; echo
cut -c 1-400 gpurun_out/${TAG}_cfg_C1.json
head -c 1500 gpurun_out/${TAG}_latency.json; echo
head -c 600 gpurun_out/${TAG}_c5s_n1.json; echo
tail -5 gpurun_out/${TAG}.err
