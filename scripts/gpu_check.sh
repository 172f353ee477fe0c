#!/bin/bash
# Runs on the GPU box under gpurun: parity tests, smoke, a short bench, and an ncu launch list.
# usage: scripts/gpu_check.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E 'passed|failed|FAILED|Error' gpurun_out/${TAG}_pytest.log | tail -40
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
tail -5 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
