#!/bin/bash
TAG=${1:-san2}
mkdir -p gpurun_out
K="small_engine or quarantin or scenario_host_rings or (random_mixed and (host or staged or 0-0)) or c1 or msg_status or capacity"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout=1000 -k "$K" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${TAG}_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/${TAG}_memcheck.log | tail -4
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout=1000 -k "small_engine or c1 or (random_mixed and 0-0) or test_broadcast_user" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/${TAG}_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|exit|Error|hazard" gpurun_out/${TAG}_racecheck.log | tail -8
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout=500 -k "small_engine or c1" > gpurun_out/${TAG}_synccheck.log 2>&1
echo "synccheck exit $?" >> gpurun_out/${TAG}_synccheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/${TAG}_synccheck.log | tail -4
