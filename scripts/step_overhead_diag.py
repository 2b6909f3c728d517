"""Where a lockstep launch's time goes: device time of srl_sim_rollout for T = 1, 2, 4, 8, 16 steps of 4096 Kuka envs (next-episode records on),
(a) in steady state with random actions and (b) right after a reset of all envs with the zero action (no contact, no joint on a limit: every
warp on the register sweep).  time(T) = fixed + T * per_step separates the launch's fixed cost (launch, state load / store, cold instruction
cache) from a step's; (a) - (b) per step is what the slowest warp's general solver path adds.  Run on the GPU box."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.backend import Backend
from srl_sim.model import load_kuka_scene

n = 4096
be = Backend(load_cuda_library(), 0)
dev = be.torch_device
st = be.stream()
blob = load_kuka_scene().blob


def fit(ts, ms):
    A = np.stack([np.ones(len(ts)), np.array(ts, float)], 1)
    (a, b), *_ = np.linalg.lstsq(A, np.array(ms), rcond=None)
    return a, b


def run(tag, prefetch, steady, zero_action):
    sim = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=blob, seed=0, prefetch_resets=prefetch)
    sim.reset(stream=st)
    Tmax = 16
    robs = be.zeros((128, n, 3), np.float32); rrew = be.zeros((128, n), np.float32); rdone = be.zeros((128, n), np.uint8)
    acts = torch.randint(0, 6, (128, n), dtype=torch.int32, device=dev)
    if steady:
        for _ in range(10):
            sim.rollout(128, acts, None, robs, rrew, rdone, None, None, stream=st)
        if prefetch:
            sim.prefetch_resets(stream=st)
    torch.cuda.synchronize()
    ts, med, mn = [1, 2, 4, 8, 16], [], []
    for T in ts:
        ms = []
        for rep in range(24):
            if not steady:
                sim.reset(stream=st)
            a = acts[(rep * Tmax) % 112:(rep * Tmax) % 112 + T].contiguous()
            if zero_action:
                a = torch.full((T, n), -1, dtype=torch.int32, device=dev)          # action -1 = no motion (the reference's "no action" of applyAction with zeros)
            sim.rollout(T, a, None, robs[:T], rrew[:T], rdone[:T], None, None, stream=st)
            ms.append(sim.last_kernel_ms())
        med.append(float(np.median(ms)) * 1e3); mn.append(float(np.min(ms)) * 1e3)
    a, b = fit(ts, med)
    print("%s: median us per launch for T = %s: %s  (min %s)  ->  fixed %.1f us + %.1f us per step"
          % (tag, ts, " ".join("%.1f" % m for m in med), " ".join("%.1f" % m for m in mn), a, b))
    sim.close()


run("steady state, random actions, records on ", True, True, False)
run("steady state, random actions, records off", False, True, False)
run("after reset, random actions, records off  ", False, False, False)
