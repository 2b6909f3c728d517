"""
Mirror of environments/registry.py:41-72 -- the ``registered_env`` table keyed by the reference's ids,
and a gym-style ``make``.  The eight PyBullet ids are kept; the real-robot / Box2D ids of the reference
(Baxter-v0, RoboboGymEnv-v0, OmnirobotEnv-v0, CarRacingGymEnv-v0) are bridges to other systems and
are out of scope of the simulator.
"""
from environments import PlottingType, ThreadingType
from environments.srl_env import SRLGymEnv
from environments.mobile_robot.mobile_robot_env import MobileRobotGymEnv
from environments.mobile_robot.mobile_robot_2target_env import MobileRobot2TargetGymEnv
from environments.mobile_robot.mobile_robot_1D_env import MobileRobot1DGymEnv
from environments.mobile_robot.mobile_robot_line_target_env import MobileRobotLineTargetGymEnv

registered_env = {
    "MobileRobotGymEnv-v0":           (MobileRobotGymEnv, SRLGymEnv, PlottingType.PLOT_2D, ThreadingType.PROCESS),
    "MobileRobot2TargetGymEnv-v0":    (MobileRobot2TargetGymEnv, MobileRobotGymEnv, PlottingType.PLOT_2D, ThreadingType.PROCESS),
    "MobileRobot1DGymEnv-v0":         (MobileRobot1DGymEnv, MobileRobotGymEnv, PlottingType.PLOT_2D, ThreadingType.PROCESS),
    "MobileRobotLineTargetGymEnv-v0": (MobileRobotLineTargetGymEnv, MobileRobotGymEnv, PlottingType.PLOT_2D, ThreadingType.PROCESS),
}

from environments.kuka_gym.kuka_button_gym_env import KukaButtonGymEnv
from environments.kuka_gym.kuka_rand_button_gym_env import KukaRandButtonGymEnv
from environments.kuka_gym.kuka_2button_gym_env import Kuka2ButtonGymEnv
from environments.kuka_gym.kuka_moving_button_gym_env import KukaMovingButtonGymEnv

registered_env.update({
    "KukaButtonGymEnv-v0":       (KukaButtonGymEnv, SRLGymEnv, PlottingType.PLOT_3D, ThreadingType.PROCESS),
    "KukaRandButtonGymEnv-v0":   (KukaRandButtonGymEnv, KukaButtonGymEnv, PlottingType.PLOT_3D, ThreadingType.PROCESS),
    "Kuka2ButtonGymEnv-v0":      (Kuka2ButtonGymEnv, KukaButtonGymEnv, PlottingType.PLOT_3D, ThreadingType.PROCESS),
    "KukaMovingButtonGymEnv-v0": (KukaMovingButtonGymEnv, KukaButtonGymEnv, PlottingType.PLOT_3D, ThreadingType.PROCESS),
})


class EnvSpec(object):
    """What the reference reads from a gym spec (environments/utils.py:10-33,60-95)."""

    def __init__(self, id_, entry_point):
        self.id = id_
        self._entry_point = entry_point
        self._kwargs = {}
        self.timestep_limit = None
        self.max_episode_steps = None
        self.reward_threshold = None
        self.tags = {}


class _Registry(object):
    def __init__(self):
        self.env_specs = {}

    def spec(self, id_):
        if id_ not in self.env_specs:
            raise KeyError("No registered env with id: {}".format(id_))
        return self.env_specs[id_]


registry = _Registry()


def register(_id, **kwargs):
    if _id in registry.env_specs:
        return
    registry.env_specs[_id] = EnvSpec(_id, kwargs["entry_point"])


for name, (env_class, _, _, _) in registered_env.items():
    register(_id=name, entry_point=env_class.__module__ + ":" + env_class.__name__,
             timestep_limit=None, reward_threshold=None)
