"""
Generate tests/golden/mobile_ref_golden.npz by running the REFERENCE MobileRobot env classes
(/root/reference/environments/mobile_robot/*.py, unmodified) with pybullet/gym stubbed out
(see _ref_stubs.py).  Run in the build container only:

    python tests/golden/gen_mobile_golden.py

Each case records, per step: action, observation (getSRLState), reward, done, robot_pos, target;
the test replays the same (seed, actions) through our env classes -> C-ABI -> kernel.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import _ref_stubs  # noqa: E402

_ref_stubs.install(os.path.join(ROOT, "robotics-rl-srl_b200"))

from environments.mobile_robot.mobile_robot_env import MobileRobotGymEnv  # noqa: E402
from environments.mobile_robot.mobile_robot_2target_env import MobileRobot2TargetGymEnv  # noqa: E402
from environments.mobile_robot.mobile_robot_1D_env import MobileRobot1DGymEnv  # noqa: E402
from environments.mobile_robot.mobile_robot_line_target_env import MobileRobotLineTargetGymEnv  # noqa: E402

assert MobileRobotGymEnv.__module__.startswith("environments.") and "/root/reference" in sys.modules[
    MobileRobotGymEnv.__module__].__file__, "must import the reference classes"

CASES = [
    # (tag, class, env kwargs, seed, episodes)
    ("base_disc", MobileRobotGymEnv, dict(is_discrete=True), 0, 2),
    ("base_disc_rand", MobileRobotGymEnv, dict(is_discrete=True, random_target=True), 3, 2),
    ("base_disc_shaped", MobileRobotGymEnv, dict(is_discrete=True, shape_reward=True, random_target=True), 11, 1),
    ("base_cont", MobileRobotGymEnv, dict(is_discrete=False), 5, 2),
    ("base_cont_rand_shaped", MobileRobotGymEnv, dict(is_discrete=False, random_target=True, shape_reward=True), 7, 1),
    ("two_target", MobileRobot2TargetGymEnv, dict(is_discrete=True), 1, 2),
    ("two_target_rand", MobileRobot2TargetGymEnv, dict(is_discrete=True, random_target=True), 2, 2),
    ("one_d", MobileRobot1DGymEnv, dict(is_discrete=True), 4, 2),
    ("one_d_rand_shaped", MobileRobot1DGymEnv, dict(is_discrete=True, random_target=True, shape_reward=True), 6, 1),
    ("line", MobileRobotLineTargetGymEnv, dict(is_discrete=True), 8, 2),
    ("line_rand_cont", MobileRobotLineTargetGymEnv, dict(is_discrete=False, random_target=True), 9, 1),
]


def run_case(cls, kwargs, seed, episodes):
    env = cls(srl_model="ground_truth", **kwargs)
    env.seed(seed)
    arng = np.random.RandomState(1000 + seed)
    rec = dict(action=[], obs=[], reward=[], done=[], robot_pos=[], target=[], reset_obs=[], reset_pos=[], reset_target=[])
    for _ in range(episodes):
        o = env.reset()
        rec["reset_obs"].append(np.asarray(o, dtype=np.float64))
        rec["reset_pos"].append(np.array(env.robot_pos, dtype=np.float64))
        rec["reset_target"].append(np.array(env.target_pos if not hasattr(env, "button_pos") else env.button_pos[0], dtype=np.float64))
        done = False
        while not done:
            if kwargs.get("is_discrete", True):
                a = int(arng.randint(env.action_space.n))
                rec["action"].append([a, 0])
            else:
                # out-of-range components exercise the clip
                a = (arng.uniform(-1.5, 1.5, size=2)).astype(np.float32)
                rec["action"].append([float(a[0]), float(a[1])])
            o, r, done, _ = env.step(a)
            rec["obs"].append(np.asarray(o, dtype=np.float64))
            rec["reward"].append(float(r))
            rec["done"].append(bool(done))
            rec["robot_pos"].append(np.array(env.robot_pos, dtype=np.float64))
            tp = env.getTargetPos()
            rec["target"].append(np.array([tp[0], tp[1] if len(tp) > 1 else 0.0], dtype=np.float64))
    return {k: np.asarray(v) for k, v in rec.items()}


def main():
    out = {}
    for tag, cls, kwargs, seed, episodes in CASES:
        rec = run_case(cls, kwargs, seed, episodes)
        for k, v in rec.items():
            out["%s/%s" % (tag, k)] = v
        print(tag, "steps:", len(rec["reward"]), "sum reward:", rec["reward"].sum())
    out["__cases__"] = np.array([c[0] for c in CASES])
    np.savez_compressed(os.path.join(HERE, "mobile_ref_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "mobile_ref_golden.npz"))


if __name__ == "__main__":
    main()
