#!/bin/bash
# full ncu captures (one launch each) of the round-2 kernels; summaries are made from the .ncu-rep files
# by scripts/ncu_summary.py in the build container
TAG=${1:-r2ncu}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:k_pack -s 6 -c 1 -o gpurun_out/${TAG}_pack python bench.py --steps 2 --warmup 3 --no-cpu --no-verify --no-secondary --no-e2e-host --sustain 0 > gpurun_out/${TAG}_pack.log 2>&1
timeout 900 $NCU -k "regex:k_offsets|k_match" -s 8 -c 2 -o gpurun_out/${TAG}_c2ctrl python bench.py --steps 2 --warmup 3 --no-cpu --no-verify --no-secondary --no-e2e-host --sustain 0 > gpurun_out/${TAG}_c2ctrl.log 2>&1
timeout 900 $NCU -k "regex:k_direct_lookup|k_pack_direct|k_offsets|k_dfill|k_dscan" -s 20 -c 5 -o gpurun_out/${TAG}_c4 python bench_configs.py --workload C4 --steps 2 --warmup 3 > gpurun_out/${TAG}_c4.log 2>&1
timeout 900 $NCU -k "regex:k_offsets|k_match|k_pack" -s 15 -c 3 -o gpurun_out/${TAG}_c5s python bench_configs.py --workload C5sparse --steps 2 --warmup 3 > gpurun_out/${TAG}_c5s.log 2>&1
timeout 900 $NCU -k "regex:k_gather_spans" -s 4 -c 1 -o gpurun_out/${TAG}_gather python bench.py --steps 2 --warmup 3 --no-cpu --no-verify --no-secondary --sustain 0 > gpurun_out/${TAG}_gather.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --sustain 0 > gpurun_out/${TAG}_launches_bench.log 2>&1
ls -la gpurun_out | grep ${TAG}
