"""
Mirror of state_representation/episode_saver.py:13-162 (``EpisodeSaver``): the dataset writer the reference envs call from
``reset()`` / ``step()`` when constructed with ``record_data=True`` (kuka_button_gym_env.py:124-126,275-276,362-363;
mobile_robot_env.py:109-111,216-217,274-275) and that ``environments.dataset_generator`` relies on.

Same files, same keys, same bookkeeping:

    <path><name>/dataset_config.json     {'relative_pos': ..., 'max_dist': '...'}
    <path><name>/env_globals.json        JSON-safe module globals of the env
    <path><name>/preprocessed_data.npz   rewards, actions, episode_starts
    <path><name>/ground_truth.npz        target_positions, ground_truth_states, images_path
    <path><name>/record_XXX/             one folder per episode

The simulator has no rasteriser, so a ``None`` observation is recorded as its frame NAME only:
``images_path`` keeps the entries the reference would have written (``<name>/record_000/frame000000``), which keeps every
array the length the reference's sanity checks (:141-146) demand; an actual image array is written with cv2 when given.
``learn_states`` (the SRL server round trip) is out of scope.
"""
import json
import os

import numpy as np


def isJsonSafe(data):
    """rl_baselines/utils.py:41-61"""
    if data is None:
        return True
    if isinstance(data, (bool, int, float, str)):
        return True
    if isinstance(data, (tuple, list)):
        return all(isJsonSafe(x) for x in data)
    if isinstance(data, dict):
        return all(isinstance(k, str) and isJsonSafe(v) for k, v in data.items())
    return False


def filterJSONSerializableObjects(input_dict):
    """rl_baselines/utils.py:64-75: sorted, JSON-safe entries only."""
    return {key: input_dict[key] for key in sorted(input_dict.keys()) if isJsonSafe(input_dict[key])}


class EpisodeSaver(object):
    """
    Save the experience data from a gym env to a file
    :param name: (str)
    :param max_dist: (float)
    :param state_dim: (int)
    :param globals_: (dict) Environments globals
    :param learn_every: (int) unused (learn_states is not supported)
    :param learn_states: (bool) must be False
    :param path: (str)
    :param relative_pos: (bool)
    """

    def __init__(self, name, max_dist, state_dim=-1, globals_=None, learn_every=3, learn_states=False,
                 path='data/', relative_pos=False):
        if learn_states:
            raise NotImplementedError("learn_states (SRL server round trip) is out of scope of the simulator")
        self.name = name
        self.data_folder = path + name
        self.path = path
        os.makedirs(self.data_folder, exist_ok=True)
        self.actions = []
        self.rewards = []
        self.target_positions = []
        self.episode_starts = []
        self.ground_truth_states = []
        self.images_path = []
        self.episode_step = 0
        self.episode_idx = -1
        self.episode_folder = None
        self.episode_success = False
        self.state_dim = state_dim
        self.n_steps = 0
        self.dataset_config = {'relative_pos': relative_pos, 'max_dist': str(max_dist)}
        with open("{}/dataset_config.json".format(self.data_folder), "w") as f:
            json.dump(self.dataset_config, f)
        if globals_ is not None:
            with open("{}/env_globals.json".format(self.data_folder), "w") as f:
                json.dump(filterJSONSerializableObjects(globals_), f)

    def saveImage(self, observation):
        """
        Record the frame (:70-88); written to disk only when an image array is supplied
        :param observation: (numpy matrix or None) BGR image
        """
        image_path = "{}/{}/frame{:06d}".format(self.data_folder, self.episode_folder, self.episode_step)
        relative_path = "{}/{}/frame{:06d}".format(self.name, self.episode_folder, self.episode_step)
        self.images_path.append(relative_path)
        if observation is not None and getattr(observation, "ndim", 0) == 3:
            import cv2
            cv2.imwrite("{}.jpg".format(image_path), cv2.cvtColor(observation[:, :, :3], cv2.COLOR_BGR2RGB))

    def reset(self, observation, target_pos, ground_truth):
        """
        Called when starting a new episode (:90-115)
        """
        if len(self.episode_starts) == 0 or self.episode_starts[-1] is False:
            self.episode_idx += 1
            self.episode_step = 0
            self.episode_success = False
            self.episode_folder = "record_{:03d}".format(self.episode_idx)
            os.makedirs("{}/{}".format(self.data_folder, self.episode_folder), exist_ok=True)
            self.episode_starts.append(True)
            self.target_positions.append(np.array(target_pos, copy=True))
            self.ground_truth_states.append(np.array(ground_truth, copy=True))
            self.saveImage(observation)

    def step(self, observation, action, reward, done, ground_truth_state):
        """
        (:117-137)
        """
        self.episode_step += 1
        self.n_steps += 1
        self.rewards.append(reward)
        self.actions.append(action)
        if reward > 0:
            self.episode_success = True
        if not done:
            self.episode_starts.append(False)
            self.ground_truth_states.append(np.array(ground_truth_state, copy=True))
            self.saveImage(observation)
        else:
            # Save the gathered data at the end of each episode
            self.save()

    def save(self):
        """
        Write data and ground truth to disk (:139-162)
        """
        assert len(self.actions) == len(self.rewards)
        assert len(self.actions) == len(self.episode_starts)
        assert len(self.actions) == len(self.images_path)
        assert len(self.actions) == len(self.ground_truth_states)
        assert len(self.target_positions) == self.episode_idx + 1
        data = {
            'rewards': np.array(self.rewards),
            'actions': np.array(self.actions),
            'episode_starts': np.array(self.episode_starts)
        }
        ground_truth = {
            'target_positions': np.array(self.target_positions),
            'ground_truth_states': np.array(self.ground_truth_states),
            'images_path': np.array(self.images_path)
        }
        np.savez('{}/preprocessed_data.npz'.format(self.data_folder), **data)
        np.savez('{}/ground_truth.npz'.format(self.data_folder), **ground_truth)
