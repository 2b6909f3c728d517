"""srl_policy_act and srl_obs_filter, 10 launches each at 4096 envs, for ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.backend import Backend
from srl_sim.policy import FusedPolicy
from rl_baselines.ppo2 import MlpPolicy, RunningNorm
n = 4096
be = Backend(load_cuda_library(), 0)
dev = be.torch_device
st = be.stream()
pol = MlpPolicy(3, n_actions=6).to(dev)
norm = RunningNorm(3, dev)
fp = FusedPolicy(be.library, pol, norm.state, seed=1)
obs = torch.randn((n, 3), device=dev)
act_env = torch.zeros(n, dtype=torch.int32, device=dev); logp = torch.zeros(n, device=dev); val = torch.zeros(n, device=dev); on = torch.zeros((n, 3), device=dev)
for _ in range(10):
    fp.filter(n, obs, on, update=True, stream=st)
    fp.act(n, on, act_env, logp, val, stream=st)
torch.cuda.synchronize()
print("done")
