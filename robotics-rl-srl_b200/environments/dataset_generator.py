"""
Mirror of ``python -m environments.dataset_generator`` (environments/dataset_generator.py:37-267): random-agent rollouts of
one registered env, recorded by ``EpisodeSaver`` -- the second consumer of the single-env API.

Same flags, same per-episode seeding (``seed = RandomState(seed).randint(1e10)``, then one distinct seed per episode,
reference :77-83,169), same output layout, same part fusion (:203-257).  Differences, all forced by what sits underneath:

* the reference forks ``--num-cpu`` OS processes, one PyBullet client each; here every env is an N=1 view on the GPU-resident
  simulator, so the ``num_cpu`` partitions run one after the other in this process and are fused exactly like the
  reference fuses its parts (the episode -> seed assignment is the reference's, so the dataset does not depend on it);
* there is no rasteriser: frames are recorded by NAME only (``images_path``), the states / targets / actions / rewards are real;
* ``--run-ppo2`` (a CnnPolicy on pixels) and ``--display`` are not available.
"""
import argparse
import glob
import os
import shutil
import time

import numpy as np

from environments.registry import registered_env


def convertImagePath(args, path, record_id_start):
    """
    Used to convert an image path, from one location, to another (reference :23-34)
    """
    image_name = path.split("/")[-1]
    new_record_id = record_id_start + int(path.split("/")[-2].split("_")[-1])
    return args.name + "/record_{:03d}".format(new_record_id) + "/" + image_name


def env_thread(args, thread_num, partition=True):
    """
    Run a session of an environment (reference :37-121, random-agent branch)
    :param args: (ArgumentParser object)
    :param thread_num: (int) The partition ID of the environment session
    :param partition: (bool) If the output should be in multiple parts (default=True)
    """
    env_kwargs = {
        "max_distance": args.max_distance,
        "random_target": args.random_target,
        "force_down": True,
        "is_discrete": not args.continuous_actions,
        "renders": False,
        "record_data": not args.no_record_data,
        "multi_view": args.multi_view,
        "save_path": args.save_path,
        "shape_reward": args.shape_reward,
        # the simulator has no rasteriser: the env records / returns the ground-truth state (the reference's default here is raw_pixels,
        # whose frames this generator only ever handed to EpisodeSaver)
        "srl_model": "ground_truth",
    }
    env_kwargs["name"] = args.name + "_part-" + str(thread_num) if partition else args.name
    env = registered_env[args.env][0](**env_kwargs)
    frames = 0
    start_time = time.time()
    # divide evenly, then do an extra one for only some of them in order to get the right count
    for i_episode in range(args.num_episode // args.num_cpu + 1 * (args.num_episode % args.num_cpu > thread_num)):
        # seed + position in this slice + size of slice (with reminder if uneven partitions)
        seed = args.seed + i_episode + args.num_episode // args.num_cpu * thread_num + \
            (thread_num if thread_num <= args.num_episode % args.num_cpu else args.num_episode % args.num_cpu)
        env.seed(seed)
        env.action_space.seed(seed)  # this is for the sample() function from gym.space
        env.reset()
        done = False
        t = 0
        while not done:
            _, _, done, _ = env.step(env.action_space.sample())
            frames += 1
            t += 1
            if done and args.verbose:
                print("Episode finished after {} timesteps".format(t + 1))
    if args.verbose:
        print("part {}: {:.2f} FPS".format(thread_num, frames / max(1e-9, time.time() - start_time)))
    env.close()
    return frames


def fuse_parts(args):
    """The reference's part fusion (:203-257)."""
    file_parts = sorted(glob.glob(args.save_path + args.name + "_part-[0-9]*"), key=lambda a: int(a.split("-")[-1]))
    os.rename(file_parts[0] + "/dataset_config.json", args.save_path + args.name + "/dataset_config.json")
    os.rename(file_parts[0] + "/env_globals.json", args.save_path + args.name + "/env_globals.json")
    ground_truth, preprocessed_data = None, None
    record_id = 0
    for part in file_parts:
        records = sorted(glob.glob(part + "/record_[0-9]*"), key=lambda a: int(a.split("_")[-1]))
        record_id_start = record_id
        for record in records:
            os.renames(record, args.save_path + args.name + "/record_{:03d}".format(record_id))
            record_id += 1
        ground_truth_load = np.load(part + "/ground_truth.npz")
        preprocessed_data_load = np.load(part + "/preprocessed_data.npz")
        gt = {arr: (np.array([convertImagePath(args, path, record_id_start) for path in ground_truth_load[arr]])
                    if arr == "images_path" else ground_truth_load[arr]) for arr in ground_truth_load.files}
        pd = {arr: preprocessed_data_load[arr] for arr in preprocessed_data_load.files}
        if ground_truth is None:
            ground_truth, preprocessed_data = gt, pd
        else:
            ground_truth = {k: np.concatenate((ground_truth[k], gt[k])) for k in gt}
            preprocessed_data = {k: np.concatenate((preprocessed_data[k], pd[k])) for k in pd}
        shutil.rmtree(part, ignore_errors=True)
    np.savez(args.save_path + args.name + "/ground_truth.npz", **ground_truth)
    np.savez(args.save_path + args.name + "/preprocessed_data.npz", **preprocessed_data)


def main(argv=None):
    parser = argparse.ArgumentParser(description='Deteministic dataset generator for SRL training ' +
                                                 '(can be used for environment testing)')
    parser.add_argument('--num-cpu', type=int, default=1, help='number of partitions (the reference: processes)')
    parser.add_argument('--num-episode', type=int, default=50, help='number of episode to run')
    parser.add_argument('--save-path', type=str, default='srl_zoo/data/', help='Folder where the environments will save the output')
    parser.add_argument('--name', type=str, default='kuka_button', help='Folder name for the output')
    parser.add_argument('--env', type=str, default='KukaButtonGymEnv-v0', help='The environment wanted', choices=list(registered_env.keys()))
    parser.add_argument('--no-record-data', action='store_true', default=False)
    parser.add_argument('--max-distance', type=float, default=0.28, help='Beyond this distance from the goal, the agent gets a negative reward')
    parser.add_argument('-c', '--continuous-actions', action='store_true', default=False)
    parser.add_argument('--seed', type=int, default=0, help='the seed')
    parser.add_argument('-f', '--force', action='store_true', default=False, help='Force the save, even if it overrides something else')
    parser.add_argument('-r', '--random-target', action='store_true', default=False, help='Set the button to a random position')
    parser.add_argument('--multi-view', action='store_true', default=False, help='accepted for compatibility (no cameras)')
    parser.add_argument('--shape-reward', action='store_true', default=False, help='Shape the reward (reward = - distance) instead of a sparse reward')
    parser.add_argument('--reward-dist', action='store_true', default=False, help='Prints out the reward distribution when the dataset generation is finished')
    parser.add_argument('--verbose', action='store_true', default=False)
    args = parser.parse_args(argv)

    assert (args.num_cpu > 0), "Error: number of cpu must be positive and non zero"
    assert (args.max_distance > 0), "Error: max distance must be positive and non zero"
    assert (args.num_episode > 0), "Error: number of episodes must be positive and non zero"
    assert not args.reward_dist or not args.shape_reward, "Error: cannot display the reward distribution for continuous reward"
    if args.num_cpu > args.num_episode:
        args.num_cpu = args.num_episode
    # this is done so seed 0 and 1 are different and not simply offset of the same datasets.
    args.seed = np.random.RandomState(args.seed).randint(int(1e10))
    if not args.no_record_data and os.path.exists(args.save_path + args.name):
        assert args.force, "Error: save directory '{}' already exists".format(args.save_path + args.name)
        shutil.rmtree(args.save_path + args.name)
        for part in glob.glob(args.save_path + args.name + "_part-[0-9]*"):
            shutil.rmtree(part)
    if not args.no_record_data:
        os.makedirs(args.save_path + args.name)
    if args.num_cpu == 1:
        frames = env_thread(args, 0, partition=False)
    else:
        frames = sum(env_thread(args, i, partition=True) for i in range(args.num_cpu))
    if not args.no_record_data and args.num_cpu > 1:
        fuse_parts(args)
    if args.reward_dist:
        rewards, counts = np.unique(np.load(args.save_path + args.name + "/preprocessed_data.npz")['rewards'], return_counts=True)
        counts = ["{:.2f}%".format(val * 100) for val in counts / np.sum(counts)]
        print("reward distribution:")
        [print(" ", reward, count) for reward, count in list(zip(rewards, counts))]
    return frames


if __name__ == '__main__':
    main()
