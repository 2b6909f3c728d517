import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from srl_sim._abi import SimLibrary, load_cuda_library
from srl_sim.backend import Backend
import test_kuka_gpu as tk
cuda = Backend(load_cuda_library(), 0); oracle = Backend(SimLibrary(os.path.join(ROOT, "oracle", "liboracle_sim.so")), -1)
n, T = 24, 900
rs = np.random.RandomState(6)
acts = rs.randint(0, 6, size=(T, n)).astype(np.int32); noise = rs.normal(0, 0.01, size=(T, n)).astype(np.float32)
cfg = dict(seed=4, random_target=True)
c = tk._run(cuda, "KukaMovingButtonGymEnv-v0", n, T, acts, noise, chunk=300, **cfg)
o = tk._run(oracle, "KukaMovingButtonGymEnv-v0", n, T, acts, noise, chunk=300, **cfg)
bad = np.argwhere(c["done"] != o["done"])
print("done mismatches", len(bad), bad[:5])
badr = np.argwhere(c["rew"] != o["rew"])
print("rew mismatches", len(badr), badr[:8])
if len(badr):
    t, i = badr[0]
    for tt in range(max(0, t - 3), min(T, t + 3)):
        print(tt, "cuda obs", c["obs"][tt, i], "rew", c["rew"][tt, i], "| oracle obs", o["obs"][tt, i], "rew", o["rew"][tt, i])
print("target diff", np.abs(c["target"] - o["target"]).max())
