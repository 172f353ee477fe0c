#!/bin/bash
# perf exploration on the GPU box: variant sweep + ncu launch list + one full capture of the top kernel
TAG=${1:-p1}
mkdir -p gpurun_out
for v in 0 4 2 1024 8 0; do
  echo "variant $v" >> gpurun_out/${TAG}_sweep.txt
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-verify --variant $v 2>>gpurun_out/${TAG}_sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('value %.1f GB/s  pack %.1f GB/s (%.3f)  ms/step %.4f  e2e %.1f  stage_ms %s' % (d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms']))" >> gpurun_out/${TAG}_sweep.txt 2>&1
done
cat gpurun_out/${TAG}_sweep.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-verify > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pack -s 4 -c 2 -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu --no-verify > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
