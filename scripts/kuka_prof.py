"""Short Kuka rollout for ncu captures: N=4096, a few warm-up rollouts then the measured ones."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.backend import Backend
from srl_sim.model import load_kuka_scene
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
epw = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cuda = Backend(load_cuda_library(), 0)
sim = cuda.make_sim("KukaButtonGymEnv-v0", N, model_blob=load_kuka_scene().blob, seed=1, envs_per_warp=epw)
sim.reset(stream=cuda.stream())
a = cuda.from_host(np.random.RandomState(1).randint(0, 6, size=(T, N)).astype(np.int32))
nz = cuda.from_host(np.random.RandomState(2).normal(0, 0.01, size=(T, N)).astype(np.float32))
obs = cuda.zeros((T, N, 3), np.float32); rew = cuda.zeros((T, N), np.float32); done = cuda.zeros((T, N), np.uint8)
for it in range(reps):
    sim.rollout(T, a, nz, obs, rew, done, stream=cuda.stream())
    print("rollout %d: %.3f ms  (%.2f M env-steps/s)" % (it, sim.last_kernel_ms(), N * T / sim.last_kernel_ms() / 1e3))
torch.cuda.synchronize()
