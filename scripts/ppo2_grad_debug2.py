import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.policy import FusedPPO2Grad
from test_policy_cpu import _policy
lib = load_cuda_library()
discrete, obs_dim, n_out, rows, mb = True, 3, 6, 5000, 4096
pol = _policy(obs_dim, discrete, n_out, seed=7).cuda()
with torch.no_grad():
    for p in pol.parameters(): p.mul_(3.0)
g = torch.Generator(device="cuda").manual_seed(5)
obs = torch.randn((rows, obs_dim), device="cuda", generator=g)
act = torch.randint(0, n_out, (rows,), device="cuda", generator=g)
with torch.no_grad(): logp0, _, v0 = pol.evaluate(obs, act)
old_logp = (logp0 + 0.4 * torch.randn(rows, device="cuda", generator=g)).contiguous()
old_val = (v0 + 0.3 * torch.randn(rows, device="cuda", generator=g)).contiguous()
adv = torch.randn(rows, device="cuda", generator=g) * 2.0 + 0.5
ret = (v0 + torch.randn(rows, device="cuda", generator=g)).contiguous()
idx = torch.randperm(rows, device="cuda", generator=g)[:mb].contiguous()
c = 0.2
# (a) torch's per-sample d vf_loss / d v against the closed form
v = pol.vf(obs[idx]).squeeze(-1); v.retain_grad()
vclip = old_val[idx] + torch.clamp(v - old_val[idx], -c, c)
loss = 0.5 * torch.max((v - ret[idx]) ** 2, (vclip - ret[idx]) ** 2).mean()
loss.backward()
gt = v.grad * mb
vd = v.detach(); ov = old_val[idx]; R = ret[idx]
dv = vd - ov; dvc = dv.clamp(-c, c); e1 = vd - R; e2 = (ov + dvc) - R; l1 = e1 * e1; l2 = e2 * e2
mine = torch.where((dvc == dv) | (l1 > l2), e1, torch.where(l1 == l2, 0.5 * e1, torch.zeros_like(e1)))
d = (gt - mine).abs()
print("closed form vs autograd on the GPU: max", float(d.max()), "n > 1e-5:", int((d > 1e-5).sum()), " sum autograd", float(gt.sum()), " sum closed form", float(mine.sum()))
# (b) the kernel with ONLY the value loss live (adv = 0 kills the policy term, ent 0), one chunk at a time: which chunk's vf.4.bias is off?
zero = torch.zeros_like(adv)
for lo in range(0, 4096, 512):
    sub = idx[lo:lo + 512].contiguous()
    for p in pol.parameters(): p.grad = None
    vv = pol.vf(obs[sub]).squeeze(-1)
    vc = old_val[sub] + torch.clamp(vv - old_val[sub], -c, c)
    (0.5 * 0.5 * torch.max((vv - ret[sub]) ** 2, (vc - ret[sub]) ** 2).mean()).backward()
    want = pol.vf[4].bias.grad.clone()
    fused = FusedPPO2Grad(lib, pol, 512)
    fused(sub, obs, act, zero, ret, old_logp, old_val, c, 0.0, 0.5, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("rows %4d..%4d  vf.4.bias got %.7e want %.7e  diff x mb/vf_coef = %.4f" % (lo, lo + 512, float(pol.vf[4].bias.grad), float(want), float(pol.vf[4].bias.grad - want) * 1024))

def kernel_vb(sub):
    fused = FusedPPO2Grad(lib, pol, int(sub.numel()))
    fused(sub.contiguous(), obs, act, zero, ret, old_logp, old_val, c, 0.0, 0.5, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return float(pol.vf[4].bias.grad) * sub.numel() / 0.5

def torch_vb(sub):
    vv = pol.vf(obs[sub]).squeeze(-1)
    dvv = vv - old_val[sub]; dcc = dvv.clamp(-c, c); a1 = vv - ret[sub]; a2 = (old_val[sub] + dcc) - ret[sub]
    m = torch.where((dcc == dvv) | (a1 * a1 > a2 * a2), a1, torch.where(a1 * a1 == a2 * a2, 0.5 * a1, torch.zeros_like(a1)))
    return float(m.sum()), vv, dvv, a1, a2

def torch_auto(sub):
    for p in pol.parameters(): p.grad = None
    vv = pol.vf(obs[sub]).squeeze(-1)
    vc = old_val[sub] + torch.clamp(vv - old_val[sub], -c, c)
    (0.5 * 0.5 * torch.max((vv - ret[sub]) ** 2, (vc - ret[sub]) ** 2).mean()).backward()
    return float(pol.vf[4].bias.grad) * sub.numel() / 0.5

sub = idx[2496:2528]
print("chunk: autograd", torch_auto(sub), " closed form", torch_vb(sub)[0], " kernel", kernel_vb(sub), " sum of kernel singles", sum(kernel_vb(sub[j:j + 1]) for j in range(32)),
      " sum of autograd singles", sum(torch_auto(sub[j:j + 1]) for j in range(32)))
for j in range(32):
    rest = torch.cat([sub[:j], sub[j + 1:]])
    k, t = kernel_vb(rest), torch_auto(rest)
    if abs(k - t) < 1e-3:
        one = sub[j:j + 1]
        pd = pol.double()
        vv64 = pd.vf(obs[one].double()).squeeze(-1)
        pol.float()
        vv = pol.vf(obs[one]).squeeze(-1)
        vv32 = pol.vf(obs[sub]).squeeze(-1)[j:j + 1]
        print("removing sample %d (row %d) fixes it: v as batch of 1 %.9g, inside the batch of 32 %.9g, float64 %.12g; old_val %.9g ret %.9g; v - old_val = %.9g (clip 0.2); e1 %.9g"
              % (j, int(one), float(vv), float(vv32), float(vv64), float(old_val[one]), float(ret[one]), float(vv - old_val[one]), float(vv - ret[one])))
