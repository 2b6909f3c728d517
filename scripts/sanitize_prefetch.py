"""Lockstep Kuka stepping with next-episode records (helper slots in every step launch + an occasional bulk fill from a side stream), small
enough for compute-sanitizer (memcheck / racecheck): short episodes so that every env consumes several records."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.backend import Backend
from srl_sim.model import load_kuka_scene
be = Backend(load_cuda_library(), 0)
n, T = 40, 60
sim = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=load_kuka_scene().blob, seed=3, max_steps=8, prefetch_resets=True, envs_per_warp=3)
main, side = torch.cuda.current_stream(), torch.cuda.Stream()
obs = be.zeros((n, 3), np.float32); rew = be.zeros((n,), np.float32); done = be.zeros((n,), np.uint8)
sim.reset(obs_out=obs, stream=main.cuda_stream)
a = be.from_host(np.random.RandomState(0).randint(0, 6, size=(T, n)).astype(np.int32))
nd = 0
for t in range(T):
    sim.step(a[t], None, obs, rew, done, None, None, stream=main.cuda_stream)
    if t % 7 == 0:
        sim.prefetch_resets(stream=side.cuda_stream)   # bulk fill from a side stream: the library orders it against the step launches
    nd += int(done.sum().item())
torch.cuda.synchronize()
print("prefetch sanitize run ok: episodes finished", nd)
