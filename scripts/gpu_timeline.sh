#!/bin/bash
# Where do the stages of pipelined batches run?  PCDN_TIMELINE dumps, per batch, the device timestamps of the stage events
# (engine.cu pcdn_release_batch).  Config 5 sparse / dense with the pack-stream overlap forced on (variant 8) and off (32).
TAG=${1:-tl}
mkdir -p gpurun_out
for wl in C5sparse C5dense; do
  for v in 8 32; do
    rm -f gpurun_out/${TAG}_${wl}_v$v.txt
    PCDN_TIMELINE=gpurun_out/${TAG}_${wl}_v$v.txt timeout 100 python bench_configs.py --workload $wl --steps 8 --warmup 3 --variant $v > gpurun_out/${TAG}_${wl}_v$v.json 2> gpurun_out/${TAG}_${wl}_v$v.err
    echo "== $wl v$v"; sed -n 4,11p gpurun_out/${TAG}_${wl}_v$v.txt
  done
done
