"""Mirror of environments/mobile_robot/mobile_robot_line_target_env.py (target is the line x = x_target)."""
from .mobile_robot_env import *  # noqa: F401,F403
from .mobile_robot_env import MobileRobotGymEnv

REWARD_DIST_THRESHOLD = 0.1
ROBOT_OFFSET = 0.2  # Take into account the robot length for computing distance to target


class MobileRobotLineTargetGymEnv(MobileRobotGymEnv):
    """
    Mobile Robot with a line target: reward when |x_target - ROBOT_OFFSET - x| <= 0.1 (reference :108-125).
    """
    _ENV_ID = "MobileRobotLineTargetGymEnv-v0"

    def __init__(self, name="mobile_robot_line_target", **kwargs):
        super(MobileRobotLineTargetGymEnv, self).__init__(name=name, **kwargs)

    def _reset_draws(self):
        # draw order of reference :50-60: only x of the target is randomised
        x_start = self._max_x / 2 + self.np_random.uniform(- self._max_x / 3, self._max_x / 3)
        y_start = self._max_y / 2 + self.np_random.uniform(- self._max_y / 3, self._max_y / 3)
        x_pos = 0.0
        if self._random_target:
            margin = 0.1 * self._max_x
            x_pos = self.np_random.uniform(self._min_x + margin, self._max_x - margin)
        return [x_start, y_start, x_pos, 0.0, 0.0, 0.0]

    def getTargetPos(self):
        # Return only the x-coordinate plus an offset to account for the robot length
        return self.target_pos[:1] - ROBOT_OFFSET
