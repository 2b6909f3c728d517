"""Mirror of environments/mobile_robot/mobile_robot_1D_env.py (1-D debug variant, 2 discrete actions)."""
import numpy as np

from .mobile_robot_env import *  # noqa: F401,F403
from .mobile_robot_env import MobileRobotGymEnv
from srl_sim import spaces

N_DISCRETE_ACTIONS = 2


class MobileRobot1DGymEnv(MobileRobotGymEnv):
    """
    1-D Mobile Robot environment: the robot moves along x only, ground truth and target are 1-vectors
    (reference :38-49).  Discrete actions only (reference :43 raises ValueError otherwise).
    """
    _ENV_ID = "MobileRobot1DGymEnv-v0"

    def __init__(self, name="mobile_robot_1D", **kwargs):
        super(MobileRobot1DGymEnv, self).__init__(name=name, **kwargs)
        self.camera_target_pos = (2, 0, 0)        # (:33)

    def _make_action_space(self):
        if self._is_discrete:
            return spaces.Discrete(N_DISCRETE_ACTIONS)
        raise ValueError("Only discrete actions is supported")

    def _reset_draws(self):
        # draw order of reference :64-72
        x_start = self._max_x / 2 + self.np_random.uniform(- self._max_x / 3, self._max_x / 3)
        x_pos = 0.0
        if self._random_target:
            margin = 0.1 * self._max_x
            x_pos = self.np_random.uniform(self._min_x + margin, self._max_x - margin)
        return [x_start, 0.0, x_pos, 0.0, 0.0, 0.0]

    def getTargetPos(self):
        # Return only the [x] coordinates
        return self.target_pos[:1]

    @staticmethod
    def getGroundTruthDim():
        return 1

    def getGroundTruth(self):
        # Return only the [x] coordinates
        return np.array(self.robot_pos)[:1]
