"""
Mirror of environments/kuka_gym/kuka_rand_button_gym_env.py: the button-push env "with a push button in a
random position and some random objects".  The reference places 10 distractor bodies and a sphere whose
choice / kick force come from the GLOBAL, unseeded ``np.random`` (:59-68,117-127), i.e. they are not
reproducible in the reference itself; the batched simulator omits them.
Everything else -- MAX_STEPS, the env RNG draws of reset() -- is identical to KukaButtonGymEnv.
"""
from .kuka_button_gym_env import *  # noqa: F401,F403
from .kuka_button_gym_env import KukaButtonGymEnv

MAX_STEPS = 1000
BALL_FORCE = 10


class KukaRandButtonGymEnv(KukaButtonGymEnv):
    """
    Kuka environment with a push button in a random position (distractor objects omitted).
    """
    _ENV_ID = "KukaRandButtonGymEnv-v0"

    def __init__(self, name="kuka_rand_button_gym", **kwargs):
        super(KukaRandButtonGymEnv, self).__init__(name=name, **kwargs)
        self.max_steps = MAX_STEPS

    def _reset_draws(self):
        # The reference consumes 2 env-RNG uniforms for each of the 10 distractor placements (:62-64) between the
        # button draws and the random init actions; keep the stream aligned.
        x_pos, y_pos = 0.5, 0
        if self._random_target:
            x_pos += 0.15 * self.np_random.uniform(-1, 1)
            y_pos += 0.3 * self.np_random.uniform(-1, 1)
        for _ in range(10):
            self.np_random.uniform(-1, 1)
            self.np_random.uniform(-1, 1)
        saved, self._random_target = self._random_target, False
        try:
            tail = super(KukaRandButtonGymEnv, self)._reset_draws()[2:]
        finally:
            self._random_target = saved
        return [x_pos, y_pos] + tail
