#!/bin/bash
# round 2, call 42: full GPU suite with the fused PPO2 gradient on by default + PPO2 timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|Error" | tail -12 ) > gpurun_out/c42_pytest.txt
( timeout 400 python scripts/ppo2_phase_timing.py 12 2>&1 | tail -3 ) > gpurun_out/c42_ppo2.txt
( timeout 300 compute-sanitizer --tool memcheck python scripts/ppo2_grad_racecheck.py 2>&1 | grep -E "ERROR SUMMARY|ran" ) > gpurun_out/c42_memcheck.txt
cat gpurun_out/c42_pytest.txt gpurun_out/c42_ppo2.txt gpurun_out/c42_memcheck.txt
