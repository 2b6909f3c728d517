#!/bin/bash
# round 2, call 27: incremental contact watch folded into the motor rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 1500 python -m pytest tests/test_kuka_gpu.py tests/test_prefetch_gpu.py -x -q -s 2>&1 | grep -E "FULL-BATCH|passed|failed|Error|assert" | tail -30 ) > gpurun_out/c27_pytest.txt
( SRL_SIM_CUDA_LIB=$PWD/robotics-rl-srl_b200/csrc/libsrl_variant_timing.so timeout 300 python scripts/lockstep_slot_timing.py 2>&1 | tail -14 ) > gpurun_out/c27_slot_timing.txt
( timeout 300 python scripts/step_launch_timing.py 2>&1 | tail -6 ) > gpurun_out/c27_step_timing.txt
( timeout 600 python -m pytest tests/test_policy_gpu.py -x -q 2>&1 | tail -3 ) > gpurun_out/c27_policy.txt
( timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/c27_bench.json
cat gpurun_out/c27_policy.txt gpurun_out/c27_pytest.txt gpurun_out/c27_slot_timing.txt gpurun_out/c27_step_timing.txt; cut -c1-400 gpurun_out/c27_bench.json
