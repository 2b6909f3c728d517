#!/bin/bash
# 2-GPU data-parallel PPO2 check (run with `gpurun --gpus 2`): torchrun, one rank per GPU, 4096 Kuka envs per rank, gradient all-reduce
# per minibatch + observation-filter merge per rollout over NCCL; prints the global env-steps/s of the training loop per update.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    -m rl_baselines.train --algo ppo2 --env KukaButtonGymEnv-v0 --num-cpu 4096 --num-timesteps 8000000 --log-dir gpurun_out/ppo2_dp2 2>&1 | tail -12 | tee gpurun_out/ppo2_dp2.txt
