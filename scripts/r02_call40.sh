#!/bin/bash
# round 2, call 40: fused PPO2 gradient -- tests, trainer tests, kernel durations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_trainer_gpu.py -x -q 2>&1 | tail -5 ) > gpurun_out/c40_test.txt
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ppo2_grad_kernel|adv_stats_kernel|ppo2_reduce_kernel" -c 12 --csv --log-file gpurun_out/c40_ppo2_kernels.csv python scripts/ppo2_phase_timing.py 2 > /dev/null 2>&1 )
( timeout 400 python scripts/ppo2_phase_timing.py 12 2>&1 | tail -3 ) > gpurun_out/c40_ppo2.txt
cat gpurun_out/c40_test.txt gpurun_out/c40_ppo2.txt; grep -E "ppo2_grad_kernel|adv_stats|ppo2_reduce" gpurun_out/c40_ppo2_kernels.csv | awk -F'","' '{print $5, $NF}' | head -12
