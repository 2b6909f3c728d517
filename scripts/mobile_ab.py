"""A/B of MobileRobot rollout launch shapes in ONE process (run on the GPU box): CTA size x library variant, config 4
(8192 envs x 1024 steps, discrete actions from HBM), cold clean L2 per launch, CUDA events on the launching stream.
Usage: python scripts/mobile_ab.py [lib.so ...]   (default: the in-tree library)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import SimLibrary, CUDA_LIBRARY_PATH
from srl_sim.backend import Backend

libs = sys.argv[1:] or [CUDA_LIBRARY_PATH]
n, T, steps = 8192, 1024, 30
acts_h = np.random.default_rng(0).integers(0, 4, (T, n), dtype=np.int32)
for lib in libs:
    be = Backend(SimLibrary(os.path.abspath(lib)), 0)
    dev = be.torch_device
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    sink = torch.zeros((), dtype=torch.float32, device=dev)
    acts = be.from_host(acts_h)
    obs = be.zeros((T, n, 2), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
    for blk in (128, 64, 32):
        os.environ["SRL_MOBILE_BLOCK"] = str(blk)
        sim = be.make_sim("MobileRobotGymEnv-v0", n, seed=0)
        st = be.stream()
        sim.reset(stream=st)
        for _ in range(5):
            sim.rollout(T, acts, None, obs, rew, done, None, None, stream=st)
        ms = []
        for k in range(steps):
            flush.fill_(k & 0xff); sink += flush_rd.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sim.rollout(T, acts, None, obs, rew, done, None, None, stream=st); e1.record()
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        ms = np.array(ms)
        gb = n * (160 + T * 17) / 1e9
        print("%s block=%3d  mean %.4f ms  min %.4f ms  %.1f G env-steps/s  %.0f GB/s (%.1f %% of 6571.6)  checksum %.6f %d"
              % (os.path.basename(lib), blk, ms.mean(), ms.min(), n * T / ms.mean() / 1e6, gb / (ms.mean() * 1e-3), 100 * gb / (ms.mean() * 1e-3) / 6571.6,
                 float(obs.double().sum().item()), int(done.sum().item())), flush=True)
        sim.close()
