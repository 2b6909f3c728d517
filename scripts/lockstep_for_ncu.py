"""A steady-state lockstep loop of 4096 Kuka envs with next-episode records, for ncu: 10 warm-up rollouts, then 40 srl_sim_step launches."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.backend import Backend
from srl_sim.model import load_kuka_scene
n, T = 4096, 128
be = Backend(load_cuda_library(), 0)
st = be.stream()
sim = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=load_kuka_scene().blob, seed=0, prefetch_resets=True)
obs = be.zeros((n, 3), np.float32); rew = be.zeros((n,), np.float32); done = be.zeros((n,), np.uint8)
sim.reset(obs_out=obs, stream=st)
acts = torch.randint(0, 6, (T, n), dtype=torch.int32, device=be.torch_device)
robs = be.zeros((T, n, 3), np.float32); rrew = be.zeros((T, n), np.float32); rdone = be.zeros((T, n), np.uint8)
for _ in range(10):
    sim.rollout(T, acts, None, robs, rrew, rdone, None, None, stream=st)
sim.prefetch_resets(stream=st); torch.cuda.synchronize()
for t in range(40):
    sim.step(acts[t], None, obs, rew, done, None, None, stream=st)
torch.cuda.synchronize()
print("done")
