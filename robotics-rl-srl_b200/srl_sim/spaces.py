"""
Minimal stand-ins for ``gym.spaces.Discrete`` / ``gym.spaces.Box`` and ``gym.Env`` (gym 0.11 API),
which the reference envs expose as ``action_space`` / ``observation_space``
(environments/kuka_gym/kuka_button_gym_env.py:149-173, environments/mobile_robot/mobile_robot_env.py:134-145).
gym is not installed in this image; only the members the reference's callers touch are provided:
``n``, ``shape``, ``dtype``, ``low``, ``high``, ``sample()``, ``seed()``, ``contains()``.
"""
import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        # hash-based like gym's ``seeding.np_random`` (srl_sim/seeding.py): the reference seeds its action spaces with values up
        # to 1e10 (environments/dataset_generator.py:82,169), which a raw ``RandomState.seed`` (32-bit) would reject
        from . import seeding
        self.np_random, seed = seeding.np_random(seed)
        return [seed]

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n):
        assert n >= 0
        self.n = int(n)
        super(Discrete, self).__init__((), np.int64)

    def sample(self):
        return self.np_random.randint(self.n)

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            as_int = int(x)
        elif isinstance(x, np.ndarray) and x.dtype.kind in "iu" and x.shape == ():
            as_int = int(x)
        else:
            return False
        return 0 <= as_int < self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n


class Box(Space):
    def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
        dtype = np.dtype(dtype)
        if shape is None:
            low = np.asarray(low)
            high = np.asarray(high)
            assert low.shape == high.shape
            shape = low.shape
        else:
            assert np.isscalar(low) and np.isscalar(high)
            low = np.full(shape, low)
            high = np.full(shape, high)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        super(Box, self).__init__(shape, dtype)

    def sample(self):
        high = self.high if self.dtype.kind == "f" else self.high.astype("int64") + 1
        return self.np_random.uniform(low=self.low, high=high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def __repr__(self):
        return "Box" + str(self.shape)

    def __eq__(self, other):
        return isinstance(other, Box) and np.allclose(self.low, other.low) and np.allclose(self.high, other.high)


class Env(object):
    """The slice of ``gym.Env`` (0.11) the reference relies on."""
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return []

    @property
    def unwrapped(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False
