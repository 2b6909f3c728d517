import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.policy import FusedPPO2Grad
from test_policy_cpu import _policy
from test_policy_gpu import _torch_ppo2_grads
lib = load_cuda_library()
for discrete, obs_dim, n_out, rows, mb in [(True, 3, 6, 5000, 4096), (True, 3, 6, 64, 32), (False, 3, 3, 3000, 1000)]:
    pol = _policy(obs_dim, discrete, n_out, seed=7).cuda()
    with torch.no_grad():
        for p in pol.parameters(): p.mul_(3.0)
    g = torch.Generator(device="cuda").manual_seed(5)
    obs = torch.randn((rows, obs_dim), device="cuda", generator=g)
    act = torch.randint(0, n_out, (rows,), device="cuda", generator=g) if discrete else torch.randn((rows, n_out), device="cuda", generator=g)
    with torch.no_grad(): logp0, _, v0 = pol.evaluate(obs, act)
    old_logp = (logp0 + 0.4 * torch.randn(rows, device="cuda", generator=g)).contiguous()
    old_val = (v0 + 0.3 * torch.randn(rows, device="cuda", generator=g)).contiguous()
    adv = torch.randn(rows, device="cuda", generator=g) * 2.0 + 0.5
    ret = (v0 + torch.randn(rows, device="cuda", generator=g)).contiguous()
    idx = torch.randperm(rows, device="cuda", generator=g)[:mb].contiguous()
    for vf_coef, ent in ((0.5, 0.01), (0.5, 0.0)):
        want = _torch_ppo2_grads(pol, idx, obs, act, adv, ret, old_logp, old_val, 0.2, ent, vf_coef)
        fused = FusedPPO2Grad(lib, pol, mb)
        fused(idx, obs, act, adv, ret, old_logp, old_val, 0.2, ent, vf_coef, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        print("case", discrete, obs_dim, n_out, rows, mb, "ent", ent)
        for (name, p), b in zip(pol.named_parameters(), want):
            d = (p.grad - b).abs()
            k = int(d.argmax())
            print("  %-12s max|err| %.3e  scale %.3e  at %d: got %.6e want %.6e   n(err > 1e-3 scale) = %d of %d"
                  % (name, float(d.max()), float(b.abs().max()), k, float(p.grad.flatten()[k]), float(b.flatten()[k]), int((d > 1e-3 * b.abs().max()).sum()), d.numel()))
