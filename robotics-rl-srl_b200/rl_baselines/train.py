"""
Mirror of the reference's training entry point ``python -m rl_baselines.train`` (rl_baselines/train.py:172-333) for the
algorithms this repo provides as consumers of the simulator: ``ppo2`` (rl_baselines/ppo2.py) and ``random_agent``
(rl_baselines/random_agent.py:28-42).  Same flag names; ``--num-cpu`` is the number of envs in the batch (per GPU when launched
with ``torchrun --nproc-per-node N -m rl_baselines.train``: data-parallel PPO2, see rl_baselines/ppo2.py).
"""
import argparse
import os
import time

from environments.registry import registered_env


def main(argv=None):
    parser = argparse.ArgumentParser(description="Train script for RL algorithms")
    parser.add_argument('--algo', default='ppo2', choices=['ppo2', 'random_agent'], type=str)
    parser.add_argument('--env', type=str, help='environment ID', default='KukaButtonGymEnv-v0', choices=list(registered_env.keys()))
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--log-dir', default='/tmp/gym/', type=str)
    parser.add_argument('--num-timesteps', type=int, default=int(1e6))
    parser.add_argument('--srl-model', type=str, default='ground_truth', choices=['ground_truth'])
    parser.add_argument('--num-cpu', help='Number of envs in the lockstep batch', type=int, default=4096)
    parser.add_argument('--action-repeat', type=int, default=1)
    parser.add_argument('--shape-reward', action='store_true', default=False)
    parser.add_argument('-c', '--continuous-actions', action='store_true', default=False)
    parser.add_argument('-r', '--random-target', action='store_true', default=False)
    parser.add_argument('--device', type=int, default=0)
    args, _ = parser.parse_known_args(argv)
    env_kwargs = dict(is_discrete=not args.continuous_actions, action_repeat=args.action_repeat, random_target=args.random_target,
                      shape_reward=args.shape_reward, srl_model=args.srl_model)
    log_dir = os.path.join(args.log_dir, args.env, args.srl_model, args.algo, time.strftime("%y-%m-%d_%Hh%M_%S"))
    num_timesteps = int(1.1 * args.num_timesteps)      # the reference trains 10 % longer (train.py:319)
    if args.algo == "ppo2":
        from rl_baselines.ppo2 import train
        from srl_sim.distributed import rank_world
        rank, world, local_rank = rank_world()
        device = args.device
        if world > 1:      # torchrun: one process per GPU, --num-cpu envs on EACH rank, gradients averaged over NCCL
            import torch
            import torch.distributed as dist
            device = local_rank
            torch.cuda.set_device(device)
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        try:
            return train(args.env, args.num_cpu, num_timesteps, seed=args.seed, env_kwargs=env_kwargs, log_dir=log_dir, device=device)
        finally:
            if world > 1 and dist.is_initialized():
                dist.destroy_process_group()
    from rl_baselines.random_agent import train
    return train(args.env, args.num_cpu, num_timesteps, seed=args.seed, env_kwargs=env_kwargs)


if __name__ == '__main__':
    main()
