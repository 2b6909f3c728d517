"""
Placeholder for environments/kuka_gym/kuka_2button_gym_env.py.  The two-button env needs a second button body and
PyBullet's null-space inverse kinematics (``use_null_space = True``, reference :79, kuka.py:34-40,147-152); neither is
implemented by the simulator yet.  The id stays registered so callers get a clear error.
"""
from .kuka_button_gym_env import *  # noqa: F401,F403
from .kuka_button_gym_env import KukaButtonGymEnv

MAX_STEPS = 1500


class Kuka2ButtonGymEnv(KukaButtonGymEnv):
    """
    Gym wrapper for Kuka environment with 2 push buttons (not implemented by the batched simulator)
    """
    _ENV_ID = "Kuka2ButtonGymEnv-v0"

    def __init__(self, name="kuka_2button_gym", max_distance=2, force_down=False, **kwargs):
        raise NotImplementedError("Kuka2ButtonGymEnv-v0 (two buttons, null-space IK) is not implemented by the batched simulator")
