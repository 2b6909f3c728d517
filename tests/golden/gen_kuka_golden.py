"""
Generate tests/golden/kuka_oracle_golden.npz: trajectories of the double-precision CPU oracle for fixed
(seed, action, noise) sequences.  PARITY UNPINNED at the PyBullet boundary (pybullet and its assets are absent;
the reference's tests pin no numbers) -- these vectors pin OUR restatement against regressions and are what the
CUDA kernels are compared with on the GPU box.  Run in the build container:

    python tests/golden/gen_kuka_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
from srl_sim import _abi  # noqa: E402
from srl_sim._abi import SimLibrary  # noqa: E402
from srl_sim.backend import Backend  # noqa: E402
from srl_sim.model import load_kuka_scene  # noqa: E402

CASES = {
    # tag: (env id, n, T, cfg)
    "button_disc": ("KukaButtonGymEnv-v0", 8, 400, dict(seed=11, is_discrete=True)),
    "button_cont_rand": ("KukaRandButtonGymEnv-v0", 8, 400, dict(seed=12, is_discrete=False, random_target=True)),
    "button_disc_shaped_rep2": ("KukaButtonGymEnv-v0", 4, 200, dict(seed=13, is_discrete=True, shape_reward=True, action_repeat=2)),
}


def inputs(tag, n, T, cfg):
    rs = np.random.RandomState(abs(hash(tag)) % (2 ** 31) if False else sum(map(ord, tag)))
    if cfg.get("is_discrete", True):
        acts = rs.randint(0, 6, size=(T, n)).astype(np.int32)
        noise = rs.normal(0, 0.01, size=(T, n)).astype(np.float32)
    else:
        acts = rs.uniform(-1, 1, size=(T, n, 3)).astype(np.float32)
        noise = rs.normal(0, 0.0001, size=(T, n)).astype(np.float32)
    return acts, noise


def run(be, env_id, n, T, cfg, acts, noise, blob):
    sim = be.make_sim(env_id, n, model_blob=blob, **cfg)
    obs0 = be.zeros((n, 3), np.float32)
    sim.reset(obs_out=obs0, stream=be.stream())
    obs = be.zeros((T, n, 3), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
    ep_ret = be.zeros((T, n), np.float32); ep_len = be.zeros((T, n), np.int32)
    sim.rollout(T, be.from_host(acts), be.from_host(noise), obs, rew, done, ep_ret, ep_len, stream=be.stream())
    out = dict(obs0=be.to_host(obs0).copy(), obs=be.to_host(obs).copy(), rew=be.to_host(rew).copy(),
               done=be.to_host(done).copy(), ep_len=be.to_host(ep_len).copy(),
               q=sim.get_state(_abi.F_JOINT_POS), qd=sim.get_state(_abi.F_JOINT_VEL), ee=sim.get_state(_abi.F_EE_POS),
               grip=sim.get_state(_abi.F_ROBOT_POS), target=sim.get_state(_abi.F_TARGET_POS))
    sim.close()
    return out


def main():
    be = Backend(SimLibrary(os.path.join(ROOT, "oracle", "liboracle_sim.so")), -1)
    blob = load_kuka_scene().blob
    out = {}
    for tag, (env_id, n, T, cfg) in CASES.items():
        acts, noise = inputs(tag, n, T, cfg)
        res = run(be, env_id, n, T, cfg, acts, noise, blob)
        for k, v in res.items():
            out["%s/%s" % (tag, k)] = v
        print(tag, "dones", int(res["done"].sum()), "reward sum", float(res["rew"].sum()))
    np.savez_compressed(os.path.join(HERE, "kuka_oracle_golden.npz"), **out)
    print("wrote kuka_oracle_golden.npz")


if __name__ == "__main__":
    main()
