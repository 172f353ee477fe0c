#!/bin/bash
# full ncu captures of the secondary kernels: k_offsets (C2) and the direct path (C4)
TAG=${1:-r1ncu2}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_offsets -s 4 -c 1 -o gpurun_out/${TAG}_offsets -f python bench.py --steps 2 --warmup 3 --no-cpu --no-verify > gpurun_out/${TAG}_offsets.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_direct_lookup|k_pack_thin|k_sort_scatter|k_sort_hist" -s 10 -c 6 -o gpurun_out/${TAG}_c4 -f python bench_configs.py --workload C4 --steps 2 --warmup 3 > gpurun_out/${TAG}_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ctrl_small" -s 20 -c 1 -o gpurun_out/${TAG}_ctrl -f python bench_configs.py --workload C1 > gpurun_out/${TAG}_ctrl.log 2>&1
ls -la gpurun_out | grep ${TAG}
