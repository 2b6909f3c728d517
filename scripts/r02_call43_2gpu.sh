#!/bin/bash
# round 2, 2-GPU check (gpurun --gpus 2): the driver's launch line for bench.py, and data-parallel PPO2 with the fused gradient
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | tail -1 ) > gpurun_out/r02_scale2_kuka.json
( timeout 400 $TR --master-port 29521 -m rl_baselines.train --algo ppo2 --env KukaButtonGymEnv-v0 --num-cpu 4096 --num-timesteps 10000000 --log-dir gpurun_out/ppo2_dp2 2>&1 | tail -6 ) > gpurun_out/r02_ppo2_dp2.txt
python -c "
import json
d=json.load(open('gpurun_out/r02_scale2_kuka.json')); print('kuka N=2', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('slowest_rank'), {k:(v.get('value')) for k,v in d.get('secondary',{}).items()})"
cat gpurun_out/r02_ppo2_dp2.txt
