#!/bin/bash
# The same timeline without blocking the host (PCDN_TIMELINE_ASYNC=1: stage events from a ring of event sets, written when the
# engine is destroyed): config 5 dense with the pack-stream overlap forced on, everything queued ahead as in the benchmark loop.
TAG=${1:-tla}
mkdir -p gpurun_out
rm -f gpurun_out/${TAG}_C5dense_v8.txt
PCDN_TIMELINE=gpurun_out/${TAG}_C5dense_v8.txt PCDN_TIMELINE_ASYNC=1 timeout 40 python bench_configs.py --workload C5dense --steps 8 --warmup 3 --variant 8 > gpurun_out/${TAG}_C5dense_v8.json 2> gpurun_out/${TAG}_C5dense_v8.err
sed -n 4,14p gpurun_out/${TAG}_C5dense_v8.txt; tail -c 300 gpurun_out/${TAG}_C5dense_v8.err
