"""
Mirror of environments/kuka_gym/kuka_moving_button_gym_env.py: the push button slides along y at BUTTON_SPEED per env
step and bounces at y = +-0.3 (reference :109-119); the target (``button_pos``) moves with it.  MAX_STEPS = 1500.
The slide + bounce run in the kernel (``SRL_ENV_KUKA_MOVING_BUTTON``); the initial direction is drawn here, first thing
in reset() like the reference (:33), and handed over with the other reset draws.
"""
from .kuka_button_gym_env import *  # noqa: F401,F403
from .kuka_button_gym_env import KukaButtonGymEnv

MAX_STEPS = 1500
BUTTON_SPEED = 0.001
BUTTON_YMIN = -0.3
BUTTON_YMAX = 0.3


class KukaMovingButtonGymEnv(KukaButtonGymEnv):
    """
    Gym wrapper for Kuka environment with a push button that is moving
    """
    _ENV_ID = "KukaMovingButtonGymEnv-v0"
    _MAX_STEPS = MAX_STEPS

    def __init__(self, name="kuka_moving_button_gym", **kwargs):
        super(KukaMovingButtonGymEnv, self).__init__(name=name, **kwargs)
        self.max_steps = MAX_STEPS
        self.button_speed = 0.0

    def _reset_draws(self):
        # random initial direction, drawn before anything else (reference :33)
        self.button_speed = BUTTON_SPEED * self.np_random.choice([-1, 1])
        draws = super(KukaMovingButtonGymEnv, self)._reset_draws()
        draws[-1] = float(self.button_speed)
        return draws
