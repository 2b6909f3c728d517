"""Small Kuka + Mobile rollouts for compute-sanitizer (memcheck): exercises reset, contact path, auto-reset."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.backend import Backend
from srl_sim.model import load_kuka_scene
be = Backend(load_cuda_library(), 0)
n, T = 24, 420
sim = be.make_sim("KukaRandButtonGymEnv-v0", n, model_blob=load_kuka_scene().blob, seed=3, random_target=True, envs_per_warp=5)
sim.reset(stream=be.stream())
rs = np.random.RandomState(0)
a = rs.randint(0, 6, size=(T, n)).astype(np.int32); a[rs.rand(T, n) < 0.4] = 4
obs = be.zeros((T, n, 3), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
sim.rollout(T, be.from_host(a), None, obs, rew, done, stream=be.stream())
torch.cuda.synchronize()
print("kuka dones", int(done.sum()), "rewards", np.unique(be.to_host(rew), return_counts=True))
m = be.make_sim("MobileRobot2TargetGymEnv-v0", 70, seed=1, random_target=True)
m.reset(stream=be.stream())
o2 = be.zeros((300, 70, 2), np.float32)
m.rollout(300, None, None, o2, None, None, stream=be.stream())
torch.cuda.synchronize()
print("mobile ok", float(o2.abs().max()))
