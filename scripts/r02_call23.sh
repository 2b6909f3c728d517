#!/bin/bash
# round 2, call 23: general solver loop on the scaled system (SAT motor rows, 128-bit row loads, 4 partial sums) + four-lane policy step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 1500 python -m pytest tests/test_kuka_gpu.py tests/test_prefetch_gpu.py -x -q -s 2>&1 | grep -E "FULL-BATCH|passed|failed|Error|assert|finished episodes found" | tail -30 ) > gpurun_out/c23_pytest.txt
( timeout 300 python scripts/step_overhead_diag.py 2>&1 ) > gpurun_out/c23_overhead.txt
( timeout 300 python scripts/step_launch_timing.py 2>&1 | tail -6 ) > gpurun_out/c23_step_timing.txt
( timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/c23_bench.json
cat gpurun_out/c23_pytest.txt gpurun_out/c23_overhead.txt gpurun_out/c23_step_timing.txt; cut -c1-400 gpurun_out/c23_bench.json
