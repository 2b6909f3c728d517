"""Exploration: CUDA vs oracle divergence on Kuka + first timings (run on the GPU box)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import SimLibrary, load_cuda_library
from srl_sim.backend import Backend
from srl_sim.model import load_kuka_scene
from srl_sim import _abi

cuda = Backend(load_cuda_library(), 0)
oracle = Backend(SimLibrary(os.path.join(ROOT, "oracle", "liboracle_sim.so")), -1)
blob = load_kuka_scene().blob


def run(be, n, T, acts, noise, chunk=None, **cfg):
    sim = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=blob, **cfg)
    obs0 = be.zeros((n, 3), np.float32)
    sim.reset(obs_out=obs0, stream=be.stream())
    obs = be.zeros((T, n, 3), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
    a = be.from_host(acts); nz = be.from_host(noise)
    qs = []
    t0 = time.time()
    if chunk:
        for s in range(0, T, chunk):
            sim.rollout(chunk, a[s:s + chunk], nz[s:s + chunk], obs[s:s + chunk], rew[s:s + chunk], done[s:s + chunk], stream=be.stream())
            qs.append(sim.get_state(_abi.F_JOINT_POS).copy())
    else:
        sim.rollout(T, a, nz, obs, rew, done, stream=be.stream())
    if be.on_gpu:
        torch.cuda.synchronize()
    dt = time.time() - t0
    return dict(obs0=be.to_host(obs0).copy(), obs=be.to_host(obs).copy(), rew=be.to_host(rew).copy(), done=be.to_host(done).copy(),
                q=sim.get_state(_abi.F_JOINT_POS), qd=sim.get_state(_abi.F_JOINT_VEL), qs=qs, dt=dt, sim=sim)


n, T = 64, 1000
rs = np.random.RandomState(0)
acts = rs.randint(0, 6, size=(T, n)).astype(np.int32)
noise = rs.normal(0, 0.01, size=(T, n)).astype(np.float32)
c = run(cuda, n, T, acts, noise, chunk=50, seed=3)
o = run(oracle, n, T, acts, noise, chunk=50, seed=3)
print("reset obs diff", np.abs(c["obs0"] - o["obs0"]).max())
same_done = (c["done"] == o["done"]); same_rew = (c["rew"] == o["rew"])
print("done mismatches", (~same_done).sum(), "rew mismatches", (~same_rew).sum(), "of", same_done.size)
# per-env first divergence
first_bad = np.full(n, T)
for i in range(n):
    bad = np.where(~(same_done[:, i] & same_rew[:, i]))[0]
    if len(bad): first_bad[i] = bad[0]
print("envs with any flag mismatch:", (first_bad < T).sum(), "first bad steps:", sorted(first_bad[first_bad < T])[:10])
dobs = np.abs(c["obs"] - o["obs"]).max(axis=2)
for i in range(n):
    dobs[first_bad[i]:, i] = 0
print("max obs diff before first flag mismatch: %.3e" % dobs.max(), " mean %.3e" % dobs.mean())
for k, (qc, qo) in enumerate(zip(c["qs"], o["qs"])):
    ok = first_bad >= (k + 1) * 50
    if k % 4 == 0:
        print("t=%4d max|dq| %.3e over %d envs" % ((k + 1) * 50, np.abs(qc - qo)[ok].max() if ok.any() else -1, ok.sum()))
print("dones cuda", c["done"].sum(), "oracle", o["done"].sum())
print("oracle time %.2fs (%.1f us/env-step)" % (o["dt"], o["dt"] / (n * T) * 1e6))

# timing
for N, epw in ((4096, 0), (4096, 4), (4096, 7), (4096, 14), (4096, 32), (16384, 0), (65536, 32)):
    T2 = 128
    sim = cuda.make_sim("KukaButtonGymEnv-v0", N, model_blob=blob, seed=1, envs_per_warp=epw)
    sim.reset(stream=cuda.stream())
    a = cuda.from_host(np.random.RandomState(1).randint(0, 6, size=(T2, N)).astype(np.int32))
    nz = cuda.from_host(np.random.RandomState(2).normal(0, 0.01, size=(T2, N)).astype(np.float32))
    obs = cuda.zeros((T2, N, 3), np.float32); rew = cuda.zeros((T2, N), np.float32); done = cuda.zeros((T2, N), np.uint8)
    ms = []
    for it in range(4):
        sim.rollout(T2, a, nz, obs, rew, done, stream=cuda.stream())
        ms.append(sim.last_kernel_ms())
    print("N=%6d epw=%2d T=%d kernel ms %s -> %.2f M env-steps/s, dones/rollout %d" % (N, epw, T2, ["%.2f" % m for m in ms], N * T2 / min(ms) / 1e3, int(done.sum())))
    sim.close()
