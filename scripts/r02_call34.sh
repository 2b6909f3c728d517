#!/bin/bash
# round 2, call 34: fused PPO2 minibatch gradient (srl_ppo2_grad)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 600 python -m pytest tests/test_policy_gpu.py -x -q -k "fused_ppo2_gradient" 2>&1 | tail -15 ) > gpurun_out/c34_test.txt
( timeout 400 python scripts/ppo2_phase_timing.py 10 2>&1 | tail -3 ) > gpurun_out/c34_ppo2.txt
cat gpurun_out/c34_test.txt gpurun_out/c34_ppo2.txt
