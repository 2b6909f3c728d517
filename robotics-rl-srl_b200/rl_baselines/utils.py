"""
Mirror of the env-construction entry point of the reference, ``rl_baselines/utils.py:194-229`` (``createEnvs``).

The reference builds ``args.num_cpu`` env thunks (``environments/utils.py:36-57``: ``makeEnv`` seeds each with
``seed + rank``) and wraps them in ``SubprocVecEnv`` / ``DummyVecEnv`` -> ``VecFrameStack`` -> ``VecNormalize``.
Here the N envs are one ``BatchedSRLVecEnv`` (one kernel launch per step); ``VecFrameStack`` / ``VecNormalize`` for the
low-dimensional ground-truth observation are small host-side wrappers with the stable-baselines 2.5 semantics
(SURVEY.md Appendix B.3): frame history zeroed on done, running mean/var, clip to +-10.
"""
import pickle

import numpy as np

from srl_sim.vec_env import BatchedSRLVecEnv


class VecFrameStack(object):
    """Concatenate the last ``n_stack`` observations along the last axis; history is zeroed when an env is done."""

    def __init__(self, venv, n_stack):
        self.venv, self.n_stack = venv, n_stack
        self.num_envs = venv.num_envs
        self.action_space = venv.action_space
        d = venv.observation_space.shape[-1]
        from srl_sim import spaces
        self.observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=(d * n_stack,), dtype=np.float32)
        self.stackedobs = np.zeros((self.num_envs, d * n_stack), np.float32)
        self._d = d

    def reset(self):
        obs = self.venv.reset()
        self.stackedobs[...] = 0
        self.stackedobs[:, -self._d:] = obs
        return self.stackedobs.copy()

    def step(self, actions):
        obs, rews, dones, infos = self.venv.step(actions)
        self.stackedobs = np.roll(self.stackedobs, shift=-self._d, axis=-1)
        self.stackedobs[dones] = 0
        self.stackedobs[:, -self._d:] = obs
        return self.stackedobs.copy(), rews, dones, infos

    def close(self):
        self.venv.close()


class RunningMeanStd(object):
    def __init__(self, shape=()):
        self.mean, self.var, self.count = np.zeros(shape, np.float64), np.ones(shape, np.float64), 1e-4

    def update(self, x):
        bm, bv, bc = x.mean(axis=0), x.var(axis=0), x.shape[0]
        delta = bm - self.mean
        tot = self.count + bc
        self.mean = self.mean + delta * bc / tot
        self.var = (self.var * self.count + bv * bc + delta ** 2 * self.count * bc / tot) / tot
        self.count = tot


class VecNormalize(object):
    """Observation normalisation (norm_obs=True, norm_reward=False as in the reference, rl_baselines/utils.py:224-227)."""

    def __init__(self, venv, training=True, norm_obs=True, norm_reward=False, clip_obs=10., epsilon=1e-8):
        self.venv, self.training, self.norm_obs, self.norm_reward = venv, training, norm_obs, norm_reward
        self.num_envs, self.action_space, self.observation_space = venv.num_envs, venv.action_space, venv.observation_space
        self.obs_rms = RunningMeanStd(venv.observation_space.shape)
        self.clip_obs, self.epsilon = clip_obs, epsilon
        self.old_obs = None

    def _obfilt(self, obs):
        if not self.norm_obs:
            return obs
        if self.training:
            self.obs_rms.update(obs)
        return np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon), -self.clip_obs, self.clip_obs).astype(np.float32)

    def reset(self):
        obs = self.venv.reset()
        self.old_obs = obs
        return self._obfilt(obs)

    def step(self, actions):
        obs, rews, dones, infos = self.venv.step(actions)
        self.old_obs = obs
        return self._obfilt(obs), rews, dones, infos

    def get_original_obs(self):
        return self.old_obs

    def save_running_average(self, path):
        with open("{}/obs_rms.pkl".format(path), "wb") as f:
            pickle.dump(self.obs_rms, f)

    def load_running_average(self, path):
        with open("{}/obs_rms.pkl".format(path), "rb") as f:
            self.obs_rms = pickle.load(f)

    def close(self):
        self.venv.close()


class DummyVecEnv(object):
    """The reference-shaped plumbing: a list of single-env objects (N = 1 views on the simulator, built by ``makeEnv`` thunks, env i
    seeded with ``seed + i``) stepped one after the other, each reset when done with the post-reset observation returned -- the worker
    semantics of stable-baselines 2.5 ``DummyVecEnv`` / ``SubprocVecEnv`` that ``createEnvs`` of the reference
    builds (rl_baselines/utils.py:216-220).  BASELINE.json configs[0] and BASELINE.md B3 are measured through it; anything that wants
    throughput uses ``BatchedSRLVecEnv`` instead."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.observation_space, self.action_space = self.envs[0].observation_space, self.envs[0].action_space

    def reset(self):
        return np.stack([env.reset() for env in self.envs])

    def step(self, actions):
        obs, rews, dones, infos = [], [], [], []
        for env, a in zip(self.envs, actions):
            o, r, d, info = env.step(a)
            if d:
                o = env.reset()
            obs.append(o); rews.append(r); dones.append(d); infos.append(info)
        return np.stack(obs), np.asarray(rews, np.float32), np.asarray(dones, bool), infos

    def close(self):
        for env in self.envs:
            env.close()


def createEnvs(args, allow_early_resets=False, env_kwargs=None, load_path_normalise=None):
    """
    :param args: (argparse.Namespace Object) needs ``env``, ``num_cpu`` (number of envs), ``seed``, ``num_stack``,
        ``srl_model``; ``device`` (optional CUDA ordinal); ``per_env_objects`` (optional, extension): build ``num_cpu`` single-env
        objects behind a ``DummyVecEnv`` like the reference does, instead of one batched env
    :param allow_early_resets: (bool) kept for signature compatibility (Monitor statistics come from the kernel)
    :param env_kwargs: (dict) The extra arguments for the environment
    :param load_path_normalise: (str) the path to loading the rolling average, None if not available or wanted.
    :return: (VecEnv-like) the batched environment, wrapped like the reference does for non-pixel observations
    """
    env_kwargs = dict(env_kwargs or {})
    env_kwargs.setdefault("srl_model", getattr(args, "srl_model", "ground_truth"))
    if getattr(args, "per_env_objects", False):
        # the reference's own shape: num_cpu env objects behind a (Dummy)VecEnv (rl_baselines/utils.py:216-220)
        from environments.utils import makeEnv
        envs = DummyVecEnv([makeEnv(args.env, args.seed, i, getattr(args, "log_dir", None), allow_early_resets=allow_early_resets, env_kwargs=env_kwargs)
                            for i in range(args.num_cpu)])
    else:
        envs = BatchedSRLVecEnv(args.env, args.num_cpu, seed=args.seed, device=getattr(args, "device", None), **env_kwargs)
    envs = VecFrameStack(envs, getattr(args, "num_stack", 1))
    if env_kwargs["srl_model"] != "raw_pixels":
        envs = VecNormalize(envs, norm_obs=True, norm_reward=False)
        if load_path_normalise is not None:
            envs.load_running_average(load_path_normalise)
    return envs


def createTensorEnvs(args, env_kwargs=None, global_env_offset=0):
    """
    The device-resident sibling of ``createEnvs`` for trainers that keep the whole collection loop on the GPU (rl_baselines/ppo2.py): the
    same arguments pick the same envs (``args.env``, ``args.num_cpu`` envs seeded ``args.seed + global index``, ``env_kwargs``), but the
    result is the bare ``BatchedSRLVecEnv`` whose ``step_tensors`` / ``sim.step`` read and write CUDA tensors -- no host round trip, hence
    no host-side ``VecFrameStack`` / ``VecNormalize`` wrappers: the trainer runs the observation filter on the device and saves it in the
    ``VecNormalize`` file format (``save_obs_rms``) so that ``createEnvs(..., load_path_normalise=...)`` can load it back.
    """
    env_kwargs = dict(env_kwargs or {})
    env_kwargs.setdefault("srl_model", getattr(args, "srl_model", "ground_truth"))
    return BatchedSRLVecEnv(args.env, args.num_cpu, seed=args.seed, device=getattr(args, "device", None), global_env_offset=global_env_offset, **env_kwargs)


def save_obs_rms(path, mean, var, count):
    """Write ``<path>/obs_rms.pkl`` -- what ``VecNormalize.save_running_average`` writes (rl_baselines/train.py:145-151 calls it when the
    model improves) -- from the moments of a device-side filter."""
    rms = RunningMeanStd(np.asarray(mean).shape)
    rms.mean, rms.var, rms.count = np.asarray(mean, np.float64).copy(), np.asarray(var, np.float64).copy(), float(count)
    with open("{}/obs_rms.pkl".format(path), "wb") as f:
        pickle.dump(rms, f)
