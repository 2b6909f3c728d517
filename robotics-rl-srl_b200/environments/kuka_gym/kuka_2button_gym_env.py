"""
Mirror of environments/kuka_gym/kuka_2button_gym_env.py: two push buttons, pressed one after the other.

Differences from the single-button env that live in the kernel (``SRL_ENV_KUKA_2BUTTON``): a second button body with its
own glider, contact with ANY link of the goal button counts (``getContactPoints`` without a link index, reference :165),
``goal_id`` bookkeeping and the two-stage reward / termination (:157-214), targets at ``Z_TABLE + BUTTON_DISTANCE_HEIGHT``
(:59,69), the large workspace box (``small_constraints=False``, :78), MAX_STEPS = 1500 and -- RECALLED pybullet 1.8.6
behaviour, see DESIGN.md section 4 -- the ``use_null_space = True`` IK call (:80, kuka.py:147-149) degenerating to the plain
damped-least-squares step with the server's default joint damping 0.5, because the four null-space lists have 7 entries
for a 14-joint body and the call passes no ``jointDamping``.
The np_random draws of reset() are made here in the reference's order, including the four it throws away (:51-57).
"""
import numpy as np

from srl_sim import _abi
from .kuka_button_gym_env import *  # noqa: F401,F403
from .kuka_button_gym_env import KukaButtonGymEnv

MAX_STEPS = 1500


class Kuka2ButtonGymEnv(KukaButtonGymEnv):
    """
    Gym wrapper for Kuka environment with 2 push buttons
    (same keyword arguments as :class:`KukaButtonGymEnv`; ``max_distance`` defaults to 2, ``force_down`` to False)
    """
    _ENV_ID = "Kuka2ButtonGymEnv-v0"
    _MAX_STEPS = MAX_STEPS

    def __init__(self, name="kuka_2button_gym", max_distance=2, force_down=False, **kwargs):
        super(Kuka2ButtonGymEnv, self).__init__(name=name, max_distance=max_distance, force_down=force_down, **kwargs)
        self.max_steps = MAX_STEPS
        self.n_contacts = [0, 0]
        self.goal_id = 0
        self.button_all_pos = []
        self.button_pressed = [False]

    def _draw_button_placement(self):
        # :49-66 -- the first button's random placement is drawn and then overwritten by `0.5 + 0.0 * uniform`, which draws twice more
        if self._random_target:
            self.np_random.uniform(-1, 1)
            self.np_random.uniform(0, 1)
        self.np_random.uniform(-1, 1)
        self.np_random.uniform(-1, 1)
        x_pos, y_pos = 0.5, -0.125
        if self._random_target:
            x_pos += 0.15 * self.np_random.uniform(-1, 1)
            y_pos += 0.175 * self.np_random.uniform(-1, 0)
        return [x_pos, y_pos]    # the SECOND button; the first one always sits at (0.5, 0.125)

    def _pull_state(self):
        super(Kuka2ButtonGymEnv, self)._pull_state()
        tb = self._sim.get_state(_abi.F_TWO_BUTTON)[0]
        self.n_contacts = [int(tb[0]), int(tb[1])]
        self.goal_id = int(tb[2])
        z = Z_TABLE + BUTTON_DISTANCE_HEIGHT
        b1 = self._sim.get_state(_abi.F_BUTTON_BASE)[0]
        self.button_all_pos = [np.array([b1[0], b1[1], z]), np.array([tb[3], tb[4], z])]
        self.button_pressed = [True] * self.goal_id + [self.n_contacts[self.goal_id] >= N_CONTACTS_BEFORE_TERMINATION]
        if self.button_pressed[-1]:
            self.button_pressed.append(False)       # the reference appends a fresh flag whenever one is set (:175)
