#!/usr/bin/env python
"""
SURVEY 8(d) configs 2 and 5: per-step divergence of the fp32 sm_100a Kuka kernel from the float64 CPU oracle on identical
(seed, action, noise) sequences -- max and percentiles of |observation difference| (gripper position relative to the button, metres)
over the env batch, per block of steps, plus the count of reward / done flag mismatches.  Run on the GPU box; writes
gpurun_out/divergence_report.txt (copied to profiles/).  The oracle is the checker here, not a product path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
from srl_sim._abi import SimLibrary, load_cuda_library   # noqa: E402
from srl_sim.backend import Backend                      # noqa: E402
from srl_sim.model import load_kuka_scene                # noqa: E402


def run(be, kind, n, T, acts, noise, **cfg):
    sim = be.make_sim(kind, n, model_blob=load_kuka_scene().blob, **cfg)
    sim.reset(stream=be.stream())
    obs = be.zeros((T, n, 3), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
    sim.rollout(T, be.from_host(acts), be.from_host(noise), obs, rew, done, None, None, stream=be.stream())
    out = be.to_host(obs).copy(), be.to_host(rew).copy(), be.to_host(done).copy()
    sim.close()
    return out


def report(name, kind, n, T, acts, noise, cfg, cuda, oracle, block, out):
    co, cr, cd = run(cuda, kind, n, T, acts, noise, **cfg)
    oo, orr, od = run(oracle, kind, n, T, acts, noise, **cfg)
    err = np.abs(co.astype(np.float64) - oo).max(axis=2)              # [T, n] metres
    # an env whose done flag differs is compared up to that step only (from there on the two run different episodes)
    mism = (cd != od) | (cr != orr)
    first = np.where(mism.any(axis=0), mism.argmax(axis=0), T)
    valid = np.arange(T)[:, None] < first[None, :]
    out.append("%s: %s, %d envs x %d steps, cfg %s" % (name, kind, n, T, cfg))
    out.append("  reward/done flag mismatches: %d of %d (env, step) pairs; envs with a mismatch: %d; episodes finished (oracle): %d"
               % (int(mism.sum()), T * n, int((first < T).sum()), int(od.sum())))
    out.append("  steps        max |d obs| [m]   p99          p90          median       (tolerance 1e-3 m)")
    for s in range(0, T, block):
        e = err[s:s + block][valid[s:s + block]]
        if e.size == 0:
            continue
        out.append("  %4d-%-4d    %.3e         %.3e    %.3e    %.3e" % (s, min(T, s + block) - 1, e.max(), np.percentile(e, 99), np.percentile(e, 90), np.median(e)))
    e = err[valid]
    out.append("  all          %.3e         %.3e    %.3e    %.3e" % (e.max(), np.percentile(e, 99), np.percentile(e, 90), np.median(e)))
    out.append("")


def main():
    cuda = Backend(load_cuda_library(), 0)
    oracle = Backend(SimLibrary(os.path.join(ROOT, "oracle", "liboracle_sim.so")), -1)
    out = []
    n, T = 128, 1000
    report("config 2", "KukaButtonGymEnv-v0", n, T, np.random.default_rng(0).integers(0, 6, (T, n), dtype=np.int32),
           np.random.default_rng(1).normal(0, 0.01, (T, n)).astype(np.float32),
           dict(seed=0, is_discrete=True, random_target=False, force_down=True, action_repeat=1, max_distance=0.8), cuda, oracle, 100, out)
    n, T = 128, 2000
    report("config 5", "KukaRandButtonGymEnv-v0", n, T, np.random.default_rng(0).uniform(-1, 1, (T, n, 3)).astype(np.float32),
           np.random.default_rng(1).normal(0, 1e-4, (T, n)).astype(np.float32), dict(seed=5, is_discrete=False, random_target=True), cuda, oracle, 200, out)
    text = "\n".join(out)
    print(text)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "divergence_report.txt"), "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
