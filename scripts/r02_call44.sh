#!/bin/bash
# round 2, call 44: PPO2 gradient kernel with 256-thread CTAs over 64-sample chunks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 600 python -m pytest tests/test_policy_gpu.py -q -k "fused_ppo2 or fused_policy_step" 2>&1 | grep -E "^E  |passed|failed" | head -10 ) > gpurun_out/c44_test.txt
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ppo2_grad_kernel" -c 4 --csv --log-file gpurun_out/c44_ppo2_kernels.csv python scripts/ppo2_phase_timing.py 2 > /dev/null 2>&1 )
( timeout 300 compute-sanitizer --tool racecheck python scripts/ppo2_grad_racecheck.py 2>&1 | grep -E "RACECHECK SUMMARY|ran" ) > gpurun_out/c44_race.txt
cat gpurun_out/c44_test.txt gpurun_out/c44_race.txt; grep -E "ppo2_grad_kernel" gpurun_out/c44_ppo2_kernels.csv | awk -F'","' '{print $5, $NF}' | head -4
