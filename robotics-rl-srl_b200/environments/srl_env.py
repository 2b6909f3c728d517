"""
``SRLGymEnv`` -- the env API boundary of the reference (environments/srl_env.py:5-102), kept
name-for-name so ``rl_baselines``-style callers consume the B200-native simulator unchanged.

Differences from the reference are confined to what sits underneath: instead of a PyBullet client
per process, a subclass owns an N=1 view (`srl_sim.Sim`, ``no_auto_reset``) on the CUDA library.
"""
from srl_sim import seeding
from srl_sim.spaces import Env


class SRLGymEnv(Env):
    metadata = {
        'render.modes': ['human', 'rgb_array'],
        'video.frames_per_second': 50
    }

    def __init__(self, *, srl_model, relative_pos, env_rank, srl_pipe):
        """
        :param srl_model: (str) The SRL_model used
        :param relative_pos: (bool) position for ground truth
        :param env_rank: (int) the number ID of the environment
        :param srl_pipe: (Queue, [Queue]) contains the input and output of the SRL model
        """
        self.env_rank = env_rank
        self.srl_pipe = srl_pipe
        self.srl_model = srl_model
        self.relative_pos = relative_pos
        self.np_random = None
        # same default as the reference: seeded with 0 until makeEnv reseeds with seed + rank
        self.seed(0)

    def getSRLState(self, observation):
        """
        :param observation: (numpy float) image (unused for ground_truth)
        :return: (numpy float)
        """
        if self.srl_model == "ground_truth":
            if self.relative_pos:
                return self.getGroundTruth() - self.getTargetPos()
            return self.getGroundTruth()
        # learned-representation path (state_representation/, out of scope): same queue protocol
        self.srl_pipe[0].put((self.env_rank, observation))
        return self.srl_pipe[1][self.env_rank].get()

    def getTargetPos(self):
        raise NotImplementedError()

    @staticmethod
    def getGroundTruthDim():
        raise NotImplementedError()

    def getGroundTruth(self):
        raise NotImplementedError()

    def seed(self, seed=None):
        """
        :param seed: (int)
        :return: ([int])
        """
        self.np_random, seed = seeding.np_random(seed)
        return [seed]

    def close(self):
        pass

    def step(self, action):
        raise NotImplementedError()

    def reset(self):
        raise NotImplementedError()

    def render(self, mode='human'):
        raise NotImplementedError()
