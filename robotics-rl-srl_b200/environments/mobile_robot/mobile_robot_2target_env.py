"""Mirror of environments/mobile_robot/mobile_robot_2target_env.py (two targets, visited in order)."""
from .mobile_robot_env import *  # noqa: F401,F403
from .mobile_robot_env import MobileRobotGymEnv

MAX_STEPS = 1500  # declared by the reference (:3) but never applied: episodes still end after 250 steps


class MobileRobot2TargetGymEnv(MobileRobotGymEnv):
    """
    Mobile Robot environment with 2 targets: the second becomes active once the first is reached
    (reference :164-185).  Discrete actions only (reference :128 raises ValueError otherwise).
    """
    _ENV_ID = "MobileRobot2TargetGymEnv-v0"

    def __init__(self, name="mobile_robot_2target", **kwargs):
        super(MobileRobot2TargetGymEnv, self).__init__(name=name, **kwargs)
        self.current_target = 0

    def _reset_draws(self):
        # draw order of reference :44-69
        x_start = self._max_x / 2 + self.np_random.uniform(- self._max_x / 3, self._max_x / 3)
        y_start = self._max_y / 2 + self.np_random.uniform(- self._max_y / 3, self._max_y / 3)
        t = [0.0, 0.0, 0.0, 0.0]
        if self._random_target:
            margin = 0.1 * self._max_x
            for k in range(2):
                t[2 * k] = self.np_random.uniform(self._min_x + margin, self._max_x - margin)
                t[2 * k + 1] = self.np_random.uniform(self._min_y + margin, self._max_y - margin)
        return [x_start, y_start] + t

    def reset(self):
        self.current_target = 0
        return super(MobileRobot2TargetGymEnv, self).reset()

    def step(self, action):
        out = super(MobileRobot2TargetGymEnv, self).step(action)
        from srl_sim import _abi
        self.current_target = int(self._sim.get_state(_abi.F_COUNTERS)[0, 0])
        return out
