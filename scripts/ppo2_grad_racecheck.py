import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.policy import FusedPPO2Grad
from test_policy_cpu import _policy
lib = load_cuda_library()
pol = _policy(3, True, 6, seed=7).cuda()
rows, mb = 200, 96
g = torch.Generator(device="cuda").manual_seed(5)
obs = torch.randn((rows, 3), device="cuda", generator=g); act = torch.randint(0, 6, (rows,), device="cuda", generator=g)
z = torch.randn(rows, device="cuda", generator=g)
fused = FusedPPO2Grad(lib, pol, mb)
fused(None, obs, act, z, z, z, z, 0.2, 0.01, 0.5, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ran")
