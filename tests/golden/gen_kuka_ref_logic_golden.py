"""
Generate tests/golden/kuka_ref_logic_golden.npz by running the REFERENCE Kuka env classes
(/root/reference/environments/kuka_gym/kuka_button_gym_env.py, kuka_rand_button_gym_env.py, kuka.py -- unmodified)
on top of tests/golden/fake_pybullet.py (the oracle's physics behind a pybullet-shaped API).  Build container only:

    python tests/golden/gen_kuka_ref_logic_golden.py

Re-pinning against the real thing (wherever `pip install pybullet==1.8.6` is possible; neither the package nor its assets
exist in the build container or on the GPU box, so this mode has never been run here):

    SRL_REFERENCE_ROOT=/path/to/robotics-rl-srl python tests/golden/gen_kuka_ref_logic_golden.py --real-pybullet

runs the SAME unmodified classes with the SAME seeds and actions on real PyBullet, writes tests/golden/kuka_pybullet_golden.npz
and prints, per case, how far the committed oracle recording is from it (first differing reward / done flag, largest gripper
position difference before it).  With that file present, tests/test_kuka_cpu.py::test_oracle_matches_real_pybullet_recording
stops skipping and checks the oracle against PyBullet at BASELINE's tolerance (1e-3 m, flags exact): that is the pin DESIGN.md
section 3 says is missing.

What this pins: every line of the reference's env-level Python (action tables, noise draws, EE clip box, the
setJointMotorControl2 gains/forces, the 500 + 5 step reset sequence, button target, reward/termination counters,
step(None), action_repeat, KukaRandButton's extra RNG draws).  What it cannot pin: PyBullet's own arithmetic.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import _ref_stubs  # noqa: E402
import fake_pybullet  # noqa: E402

REAL_PYBULLET = "--real-pybullet" in sys.argv
_ref_stubs.install(os.path.join(ROOT, "robotics-rl-srl_b200"), real_pybullet=REAL_PYBULLET)
if not REAL_PYBULLET:
    sys.modules["pybullet"] = fake_pybullet.as_module()      # replace the inert stub by the oracle-backed one
import torch  # noqa: E402,F401  (the reference imports it)

from environments.kuka_gym.kuka_button_gym_env import KukaButtonGymEnv  # noqa: E402
from environments.kuka_gym.kuka_rand_button_gym_env import KukaRandButtonGymEnv  # noqa: E402
from environments.kuka_gym.kuka_moving_button_gym_env import KukaMovingButtonGymEnv  # noqa: E402
from environments.kuka_gym.kuka_2button_gym_env import Kuka2ButtonGymEnv  # noqa: E402

assert _ref_stubs.REFERENCE_ROOT in sys.modules[KukaButtonGymEnv.__module__].__file__, "must import the reference classes"

CASES = [
    # tag, class name, kwargs, seed, max env steps recorded
    ("disc", "KukaButtonGymEnv", dict(is_discrete=True), 0, 700),
    ("disc_rand_shaped", "KukaButtonGymEnv", dict(is_discrete=True, random_target=True, shape_reward=True), 1, 500),
    ("cont", "KukaButtonGymEnv", dict(is_discrete=False), 2, 400),
    ("cont_shaped_up", "KukaButtonGymEnv", dict(is_discrete=False, shape_reward=True, force_down=False), 3, 400),
    ("disc_rep3_none", "KukaButtonGymEnv", dict(is_discrete=True, action_repeat=3), 4, 300),
    ("rand_button_cont", "KukaRandButtonGymEnv", dict(is_discrete=False, random_target=True), 5, 400),
    ("rand_button_disc", "KukaRandButtonGymEnv", dict(is_discrete=True, random_target=True), 6, 500),
    ("moving_disc", "KukaMovingButtonGymEnv", dict(is_discrete=True), 7, 500),
    ("moving_cont_rand", "KukaMovingButtonGymEnv", dict(is_discrete=False, random_target=True), 8, 400),
    ("joints", "KukaButtonGymEnv", dict(is_discrete=False, action_joints=True), 9, 400),
    ("joints_shaped_none", "KukaButtonGymEnv", dict(is_discrete=False, action_joints=True, shape_reward=True), 10, 300),
    ("two_disc", "Kuka2ButtonGymEnv", dict(is_discrete=True, force_down=True), 11, 900),
    ("two_disc_rand_shaped", "Kuka2ButtonGymEnv", dict(is_discrete=True, random_target=True, shape_reward=True, force_down=True), 12, 900),
    ("two_cont_up", "Kuka2ButtonGymEnv", dict(is_discrete=False), 13, 500),
]
CLASSES = {"KukaButtonGymEnv": KukaButtonGymEnv, "KukaRandButtonGymEnv": KukaRandButtonGymEnv,
           "KukaMovingButtonGymEnv": KukaMovingButtonGymEnv, "Kuka2ButtonGymEnv": Kuka2ButtonGymEnv}


def make_actions(tag, kwargs, n, seed):
    rs = np.random.RandomState(500 + seed)
    if kwargs.get("is_discrete", True):
        a = rs.randint(0, 6, size=n).astype(np.float64)
        # a biased descent makes the episodes end through contacts as well as through the step limit
        a[rs.rand(n) < 0.35] = 4
        if "none" in tag:
            a[rs.rand(n) < 0.1] = -1           # -1 encodes step(None)
        return np.stack([a, np.zeros(n), np.zeros(n)], axis=1)
    if kwargs.get("action_joints", False):
        # joint-space set-points: a slowly varying random walk bends the arm towards the table for part of the episode
        a = np.clip(np.cumsum(rs.normal(0, 0.15, size=(n, 7)), axis=0), -1, 1)
        a[:, 1] = np.clip(a[:, 1] + np.linspace(0, 4, n) % 2.0, -1, 1)
        if "none" in tag:
            a[rs.rand(n) < 0.1, 0] = np.nan      # NaN in column 0 encodes step(None)
        return a.astype(np.float32).astype(np.float64)
    a = rs.uniform(-1, 1, size=(n, 3))
    a[:, 2] = -np.abs(a[:, 2]) if "up" not in tag else a[:, 2]
    return a.astype(np.float32).astype(np.float64)


def run_case(tag, clsname, kwargs, seed, nsteps):
    env = CLASSES[clsname](srl_model="ground_truth", **kwargs)
    env.seed(seed)
    np.random.seed(1234)                         # the reference's KukaRandButton also draws from the GLOBAL numpy RNG
    actions = make_actions(tag, kwargs, nsteps, seed)
    prs = np.random.RandomState(900 + seed)
    rec = dict(action=actions, obs=[], reward=[], done=[], arm=[], target=[], reset_obs=[], reset_target=[], reset_at=[])
    t = 0
    while t < nsteps:
        o = env.reset()
        rec["reset_obs"].append(np.asarray(o, np.float64)); rec["reset_target"].append(np.array(env.getTargetPos(), dtype=np.float64, copy=True))
        rec["reset_at"].append(t)
        done = False
        while not done and t < nsteps:
            if tag.startswith("two_") and prs.rand() < 0.85:
                # two-button cases: a greedy controller on the observation (arm - goal) presses the buttons one after the
                # other, so that goal switching / the second-button termination are exercised; the chosen action is recorded
                ob = np.asarray(o, np.float64)
                if kwargs.get("is_discrete", True):
                    if max(abs(ob[0]), abs(ob[1])) < 0.02:
                        actions[t, 0] = 4
                    elif abs(ob[0]) > abs(ob[1]):
                        actions[t, 0] = 0 if ob[0] > 0 else 1
                    else:
                        actions[t, 0] = 2 if ob[1] > 0 else 3
                else:
                    d = -ob / max(1e-9, np.abs(ob[:2]).max())
                    actions[t] = np.clip([d[0], d[1], -1.0 if np.abs(ob[:2]).max() < 0.02 else 0.0], -1, 1).astype(np.float32)
            if kwargs.get("is_discrete", True):
                a = None if actions[t, 0] < 0 else int(actions[t, 0])
            else:
                a = None if np.isnan(actions[t, 0]) else actions[t].astype(np.float32)
            o, r, done, _ = env.step(a)
            rec["obs"].append(np.asarray(o, np.float64)); rec["reward"].append(float(r)); rec["done"].append(bool(done))
            rec["arm"].append(np.array(env.getArmPos(), dtype=np.float64, copy=True)); rec["target"].append(np.array(env.getTargetPos(), dtype=np.float64, copy=True))
            if tag.startswith("two_"):
                rec.setdefault("goal", []).append(int(env.goal_id)); rec.setdefault("ncontacts", []).append(list(env.n_contacts))
            t += 1
    return {k: np.asarray(v) for k, v in rec.items()}


def compare_with_oracle_recording(out):
    """--real-pybullet: distance of the committed oracle recording (same classes, seeds, actions) from the PyBullet one."""
    ref = np.load(os.path.join(HERE, "kuka_ref_logic_golden.npz"))
    for tag, _, _, _, _ in CASES:
        if tag.startswith("two_"):
            continue                                   # closed-loop controller: the action sequences themselves diverge
        n = min(len(out[tag + "/reward"]), len(ref[tag + "/reward"]))
        flags_differ = (out[tag + "/reward"][:n] != ref[tag + "/reward"][:n]) | (out[tag + "/done"][:n] != ref[tag + "/done"][:n])
        first = int(np.argmax(flags_differ)) if flags_differ.any() else n
        d_arm = np.abs(out[tag + "/arm"][:first] - ref[tag + "/arm"][:first]).max() if first else float("nan")
        print("%-22s first differing reward/done flag at step %d of %d; max |gripper position difference| before it %.3e m" % (tag, first, n, d_arm))


def main():
    out = {}
    for tag, clsname, kwargs, seed, nsteps in CASES:
        rec = run_case(tag, clsname, kwargs, seed, nsteps)
        for k, v in rec.items():
            out["%s/%s" % (tag, k)] = v
        print(tag, "steps", len(rec["reward"]), "episodes", len(rec["reset_at"]), "dones", int(rec["done"].sum()), "reward sum %.3f" % rec["reward"].sum())
    name = "kuka_pybullet_golden.npz" if REAL_PYBULLET else "kuka_ref_logic_golden.npz"
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote " + name)
    if REAL_PYBULLET:
        compare_with_oracle_recording(out)


if __name__ == "__main__":
    main()
