#!/bin/bash
# round 2, call 22: srl_policy_act over 4 lanes per env (128 CTAs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py -x -q > gpurun_out/c22_policy.txt 2>&1; echo "policy rc=$?" >> gpurun_out/c22_policy.txt
timeout 600 python scripts/step_launch_timing.py > gpurun_out/c22_step_timing.txt 2>&1
tail -4 gpurun_out/c22_policy.txt; cat gpurun_out/c22_step_timing.txt
