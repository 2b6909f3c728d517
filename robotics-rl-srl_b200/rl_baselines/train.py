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

# rl_baselines/rl_algorithm/ppo2.py:25-36 (getOptParam): the hyper-parameters `--hyperparam name:value` may set, and their types
PPO2_OPT_PARAM = {"lam": float, "gamma": float, "max_grad_norm": float, "vf_coef": float, "learning_rate": float, "ent_coef": float,
                  "cliprange": float, "noptepochs": int, "n_steps": int}


def parserHyperParam(pairs):
    """``["name:value", ...]`` -> typed dict (train.py:321 + base_classes.py:62-80: unknown names are an AssertionError)."""
    parsed = {}
    for param in pairs:
        name, val = param.split(":")[0], param.split(":")[1]
        if name not in PPO2_OPT_PARAM:
            raise AssertionError("Error: hyperparameter {} not in list of valid hyperparameters".format(name))
        parsed[name] = PPO2_OPT_PARAM[name](val)
    return parsed


def main(argv=None):
    parser = argparse.ArgumentParser(description="Train script for RL algorithms")
    parser.add_argument('--algo', default='ppo2', choices=['ppo2', 'random_agent'], type=str)
    parser.add_argument('--env', type=str, help='environment ID', default='KukaButtonGymEnv-v0', choices=list(registered_env.keys()))
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--episode_window', type=int, default=40, help='Episode window for moving average plot (default: 40)')
    parser.add_argument('--num-stack', type=int, default=1, help='number of frames to stack (default: 1; state observations are not stacked here)')
    parser.add_argument('-joints', '--action-joints', action='store_true', default=False, help='set actions to the joints of the arm directly')
    parser.add_argument('--hyperparam', type=str, nargs='+', default=[], help='PPO2 hyper-parameters as name:value pairs')
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
    # sanity checks of the reference (train.py:221-224,265-266)
    assert args.episode_window >= 1, "Error: --episode_window cannot be less than 1"
    assert args.num_timesteps >= 1, "Error: --num-timesteps cannot be less than 1"
    assert args.num_stack == 1, "Error: --num-stack > 1 is for image observations, which this simulator does not render"
    assert args.action_repeat >= 1, "Error: --action-repeat cannot be less than 1"
    if args.action_joints and not args.continuous_actions:
        raise ValueError("The joints action space is continuous only: use '-joints' together with '-c' (kuka_button_gym_env.py:149-161)")
    hyperparams = parserHyperParam(args.hyperparam)
    env_kwargs = dict(is_discrete=not args.continuous_actions, action_repeat=args.action_repeat, random_target=args.random_target,
                      shape_reward=args.shape_reward, srl_model=args.srl_model)
    if args.action_joints:
        env_kwargs["action_joints"] = True
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
            return train(args.env, args.num_cpu, num_timesteps, seed=args.seed, env_kwargs=env_kwargs, log_dir=log_dir, device=device,
                         hyperparams=hyperparams, episode_window=args.episode_window)
        finally:
            if world > 1 and dist.is_initialized():
                dist.destroy_process_group()
    from rl_baselines.random_agent import train
    return train(args.env, args.num_cpu, num_timesteps, seed=args.seed, env_kwargs=env_kwargs)


if __name__ == '__main__':
    main()
