#!/bin/bash
# final validation of a round on one B200: parity tests, smoke, bench (with CPU baseline), launch list,
# one full ncu capture of k_pack, the host-ring egress mode
TAG=${1:-r1final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err
cut -c 1-2500 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cut -c 1-600 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py --conns 65536 --host-rings --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_hostrings.json 2>> gpurun_out/${TAG}_bench.err
cut -c 1-700 gpurun_out/${TAG}_bench_hostrings.json
for wl in C3 C4 C5dense C5sparse; do
  timeout 300 python bench_configs.py --workload $wl --steps 20 --warmup 3 > gpurun_out/${TAG}_cfg_${wl}.json 2>> gpurun_out/${TAG}_bench.err
  python -c "
import json
d = json.loads(open('gpurun_out/${TAG}_cfg_${wl}.json').readline()); print('$wl', round(d['value'], 1), round(d['ms_per_step'], 4), d['roofline']['stage_ms'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-verify > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pack -s 4 -c 2 -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu --no-verify > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
