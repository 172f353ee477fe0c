#!/bin/bash
TAG=${1:-it3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q -x --timeout=800 -k "c4 or c1 or quarantin or hot_recipient or wrap or c5" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${TAG}_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/${TAG}_memcheck.log | tail -5
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_configs.py -m gpu -q -x --timeout=800 -k "c4 or c5" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/${TAG}_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|exit" gpurun_out/${TAG}_racecheck.log | tail -5
