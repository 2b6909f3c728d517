#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 600 python -m pytest tests/test_policy_gpu.py -q -k "fused_ppo2_gradient" 2>&1 | grep -E "^E  |passed|failed" | head -20 ) > gpurun_out/c41_test.txt
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ppo2_grad_kernel|adv_stats_kernel|ppo2_reduce_kernel" -c 6 --csv --log-file gpurun_out/c41_ppo2_kernels.csv python scripts/ppo2_phase_timing.py 2 > /dev/null 2>&1 )
cat gpurun_out/c41_test.txt; grep -E "ppo2_grad_kernel|adv_stats|ppo2_reduce" gpurun_out/c41_ppo2_kernels.csv | awk -F'","' '{print $5, $NF}' | head -6
