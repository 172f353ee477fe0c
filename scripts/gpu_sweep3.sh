#!/bin/bash
TAG=${1:-p6}
mkdir -p gpurun_out
for v in 0 0 4 1024 8; do
  echo "variant $v" >> gpurun_out/${TAG}_sweep.txt
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-verify --variant $v 2>>gpurun_out/${TAG}_sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('value %.1f GB/s  pack %.1f GB/s (%.3f)  ms/step %.4f  e2e %.1f' % (d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))" >> gpurun_out/${TAG}_sweep.txt 2>&1
done
cat gpurun_out/${TAG}_sweep.txt
