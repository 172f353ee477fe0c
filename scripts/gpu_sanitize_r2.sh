#!/bin/bash
# compute-sanitizer over the paths that are new in round 2: sort-free direct path (hot recipient, c4), run-length
# span tables, output pool (look-back in the fused kernel, k_pool_finish, back-pressure + retry), egress gather,
# shards on one GPU
TAG=${1:-r2san}
mkdir -p gpurun_out
K="hot_recipient or c4 or span_runs or output_pool or (random_mixed and (0-pool or 0-runs or 0-shards-host or 0-0)) or (drain_to_host and (pool or hbm-small)) or sharded_device_resident"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout=1400 -k "$K" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${TAG}_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/${TAG}_memcheck.log | tail -4
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout=1400 -k "span_runs or output_pool or (random_mixed and (0-pool-staged-runs or 0-runs)) or test_direct_user_to_user" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/${TAG}_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|exit|Error|hazard" gpurun_out/${TAG}_racecheck.log | tail -8
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout=800 -k "span_runs or output_pool" > gpurun_out/${TAG}_synccheck.log 2>&1
echo "synccheck exit $?" >> gpurun_out/${TAG}_synccheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/${TAG}_synccheck.log | tail -4
