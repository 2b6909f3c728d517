"""Diagnose the full-batch no_force_down case: where do CUDA (four lanes per env / one thread per env) and the oracle separate?"""
import os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from srl_sim._abi import SimLibrary, load_cuda_library
from srl_sim.backend import Backend
import test_kuka_gpu as tk
env_id, T, kind, cfg, noise_std, bound = tk.FULL_BATCH_CASES["no_force_down"]
n = 4096
acts = np.random.default_rng(0).integers(0, 6, (T, n), dtype=np.int32)
noise = np.random.default_rng(1).normal(0, noise_std, (T, n)).astype(np.float32)
cuda = Backend(load_cuda_library(), 0)
res = {}
for name, epw in (("coop", 0), ("thread", 32)):
    res[name] = tk._run(cuda, env_id, n, T, acts, noise, envs_per_warp=epw, **cfg)
ob = Backend(SimLibrary(os.path.join(ROOT, "oracle", "liboracle_sim.so")), -1)
threads = 16; bounds = np.linspace(0, n, threads + 1).astype(int); parts = [None] * threads
def work(k):
    lo, hi = int(bounds[k]), int(bounds[k + 1])
    parts[k] = tk._run(ob, env_id, hi - lo, T, np.ascontiguousarray(acts[:, lo:hi]), np.ascontiguousarray(noise[:, lo:hi]), global_env_offset=lo, **cfg)
ths = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
[t.start() for t in ths]; [t.join() for t in ths]
o = {k: np.concatenate([p[k] for p in parts], axis=1) for k in ("obs", "rew", "done")}
for name in ("coop", "thread"):
    c = res[name]
    worst = []
    for i in range(n):
        bad = np.nonzero((c["rew"][:, i] != o["rew"][:, i]) | (c["done"][:, i] != o["done"][:, i]))[0]
        t_end = T if len(bad) == 0 else int(bad[0])
        d = np.abs(c["obs"][:t_end, i] - o["obs"][:t_end, i]).max(axis=1) if t_end else np.zeros(0)
        if len(d) and d.max() > 2e-4:
            worst.append((float(d.max()), i, int(d.argmax()), t_end))
    worst.sort(reverse=True)
    print(name, "envs with |obs diff| > 0.2 mm before their first flag difference:", len(worst), "worst (diff, env, step, t_end):", worst[:6])
    for dmax, i, t, t_end in worst[:2]:
        d = np.abs(c["obs"][:t_end, i] - o["obs"][:t_end, i]).max(axis=1)
        first = int(np.nonzero(d > 1e-4)[0][0])
        print("  env", i, "first step above 0.1 mm:", first, "diffs around it:", np.round(d[max(0, first - 3):first + 8] * 1e3, 3), "mm; rewards", o["rew"][max(0, first - 3):first + 4, i],
              "dones before:", int(o["done"][:first, i].sum()), "obs", o["obs"][first, i])
cc, ct = res["coop"], res["thread"]
print("coop vs thread-per-env: max |obs diff| %.3g, flags identical: %s" % (np.abs(cc["obs"] - ct["obs"]).max(), np.array_equal(cc["rew"], ct["rew"]) and np.array_equal(cc["done"], ct["done"])))
