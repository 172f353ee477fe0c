#!/bin/bash
# secondary-config sweep: CTAs/SM (bits 8+) and message-major tile size (bits 4-7) on C4 / C5sparse / C3
TAG=${1:-p7}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_sweep.txt
for wl in C4 C5sparse C3; do
  for v in 0 1536 2048 16 32 48 64 1568 1584; do
    echo -n "$wl variant $v: " >> $OUT
    timeout 300 python bench_configs.py --workload $wl --steps 10 --warmup 3 --variant $v 2>>gpurun_out/${TAG}_sweep.err | python -c "
import sys, json
for line in sys.stdin:
    if not line.startswith('{'): continue
    d = json.loads(line)
    print('value %.1f  ms/step %.4f  pack %.1f (%.3f) stage %s' % (d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['stage_ms']))
    break" >> $OUT 2>&1
  done
done
cat $OUT
