"""What one LOCKSTEP env step costs on the device (config 3's collection loop is 128 of them per update): device time of srl_sim_step for
4096 Kuka envs in steady state -- per launch, with the share of launches in which at least one env finishes an episode -- next to the fused
128-step rollout, and of the two consumer kernels (srl_policy_act, srl_obs_filter).  Run on the GPU box."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.backend import Backend
from srl_sim.model import load_kuka_scene
from srl_sim.policy import FusedPolicy
from rl_baselines.ppo2 import MlpPolicy, RunningNorm

n, T = 4096, 128
be = Backend(load_cuda_library(), 0)
dev = be.torch_device
sim = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=load_kuka_scene().blob, seed=0)
st = be.stream()
obs = be.zeros((n, 3), np.float32); rew = be.zeros((n,), np.float32); done = be.zeros((n,), np.uint8)
sim.reset(obs_out=obs, stream=st)
acts = torch.randint(0, 6, (T, n), dtype=torch.int32, device=dev)
robs = be.zeros((T, n, 3), np.float32); rrew = be.zeros((T, n), np.float32); rdone = be.zeros((T, n), np.uint8)
for _ in range(10):                       # 1280 steps: episodes are de-synchronised (steady state)
    sim.rollout(T, acts, None, robs, rrew, rdone, None, None, stream=st)
torch.cuda.synchronize()
roll_ms = []
for _ in range(3):
    sim.rollout(T, acts, None, robs, rrew, rdone, None, None, stream=st); roll_ms.append(sim.last_kernel_ms())
step_ms, any_done = [], []
for t in range(T):
    sim.step(acts[t], None, obs, rew, done, None, None, stream=st)
    step_ms.append(sim.last_kernel_ms()); any_done.append(int(done.sum().item()))
step_ms, any_done = np.array(step_ms), np.array(any_done)
print("fused rollout: %.3f ms per %d steps = %.1f us per env step" % (np.mean(roll_ms), T, 1e3 * np.mean(roll_ms) / T))
print("lockstep srl_sim_step: mean %.1f us, median %.1f us, min %.1f us, max %.1f us per launch; launches with >= 1 finished episode: %d of %d"
      % (1e3 * step_ms.mean(), 1e3 * np.median(step_ms), 1e3 * step_ms.min(), 1e3 * step_ms.max(), int((any_done > 0).sum()), T))
if (any_done > 0).any() and (any_done == 0).any():
    print("  launches with a finished episode: %.1f us; without: %.1f us" % (1e3 * step_ms[any_done > 0].mean(), 1e3 * step_ms[any_done == 0].mean()))
# next-episode records (srl_cfg.prefetch_resets): the same lockstep loop; the helper slots of every launch keep the records up
sim2 = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=load_kuka_scene().blob, seed=0, prefetch_resets=True)
sim2.reset(obs_out=obs, stream=st)
for _ in range(10):
    sim2.rollout(T, acts, None, robs, rrew, rdone, None, None, stream=st)
sim2.prefetch_resets(stream=st); torch.cuda.synchronize()
ms2, hits, fin = [], 0, 0
from srl_sim import _abi
for rep in range(3):
    for t in range(T):
        if rep == 2:
            rec = sim2.get_state(_abi.F_NEXT_RECORD); live = sim2.get_state(_abi.F_COUNTERS)[:, 3]
            ready = (rec[:, 0] == 1) & (rec[:, 2] == live)
        sim2.step(acts[t], None, obs, rew, done, None, None, stream=st)
        if rep == 2:
            ms2.append(sim2.last_kernel_ms())
            d = done.cpu().numpy().astype(bool); hits += int((d & ready).sum()); fin += int(d.sum())
torch.cuda.synchronize()
ms2 = np.array(ms2)
print("lockstep srl_sim_step WITH next-episode records (helper slots): mean %.1f us, median %.1f us, min %.1f us, max %.1f us per launch; %d of %d finished episodes took a record"
      % (1e3 * ms2.mean(), 1e3 * np.median(ms2), 1e3 * ms2.min(), 1e3 * ms2.max(), hits, fin))
pol = MlpPolicy(3, n_actions=6).to(dev)
norm = RunningNorm(3, dev)
fp = FusedPolicy(be.library, pol, norm.state, seed=1)
act_env = torch.zeros(n, dtype=torch.int32, device=dev); logp = torch.zeros(n, device=dev); val = torch.zeros(n, device=dev); on = torch.zeros((n, 3), device=dev)
for name, fn in (("srl_policy_act", lambda: fp.act(n, on, act_env, logp, val, stream=st)), ("srl_obs_filter", lambda: fp.filter(n, obs, on, update=True, stream=st))):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%s: %.1f us per launch (50 back-to-back launches, %d envs)" % (name, 1e3 * e0.elapsed_time(e1) / 50, n))
    # the same 50 launches replayed from a captured CUDA graph: device time without the host's enqueue rate (what the trainer's captured
    # collection loop pays)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(50):
                fn_s = (lambda: fp.act(n, on, act_env, logp, val, stream=side.cuda_stream)) if name == "srl_policy_act" else (lambda: fp.filter(n, obs, on, update=True, stream=side.cuda_stream))
                fn_s()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side); g.replay(); e1.record(side); torch.cuda.synchronize()
    print("%s: %.1f us per launch inside a captured graph of 50" % (name, 1e3 * e0.elapsed_time(e1) / 50))
