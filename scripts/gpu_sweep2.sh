#!/bin/bash
TAG=${1:-p2}
mkdir -p gpurun_out
python scripts/gpu_micro.py > gpurun_out/${TAG}_micro.json 2>&1
cat gpurun_out/${TAG}_micro.json
run() { echo "== $*" >> gpurun_out/${TAG}_sweep.txt; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-verify "$@" 2>>gpurun_out/${TAG}_sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('value %.1f GB/s  pack %.1f GB/s (%.3f)  ms/step %.4f  e2e %.1f' % (d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))" >> gpurun_out/${TAG}_sweep.txt 2>&1; }
run --payload 256 --conns 1048576
run --payload 1024
run --payload 1024 --ring-records 17
run --payload 1024 --ring-records 24
run --payload 1024 --ring-records 8
run --payload 4096 --conns 524288
run --payload 16000 --conns 131072
run --payload 1024 --msgs 1
run --payload 1024 --msgs 16 --ring-records 32 --conns 524288
cat gpurun_out/${TAG}_sweep.txt
