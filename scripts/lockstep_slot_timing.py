"""Which env slots make a lockstep launch long: a -DKK_TIMING build (scripts/build_variant.sh timing -DKK_TIMING, SRL_SIM_CUDA_LIB=...) records, per
env slot of the last launch, the cycles from kernel entry to the slot's exit and what its physics step did.  Steady state, 4096 Kuka envs,
next-episode records on; 60 lockstep launches; prints the slowest slots' flags and the mean exit time per category."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.backend import Backend
from srl_sim.model import load_kuka_scene
n, T = 4096, 128
lib = load_cuda_library()
be = Backend(lib, 0)
st = be.stream()
sim = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=load_kuka_scene().blob, seed=0, prefetch_resets=True)
obs = be.zeros((n, 3), np.float32); rew = be.zeros((n,), np.float32); done = be.zeros((n,), np.uint8)
sim.reset(obs_out=obs, stream=st)
acts = torch.randint(0, 6, (T, n), dtype=torch.int32, device=be.torch_device)
robs = be.zeros((T, n, 3), np.float32); rrew = be.zeros((T, n), np.float32); rdone = be.zeros((T, n), np.uint8)
for _ in range(10):
    sim.rollout(T, acts, None, robs, rrew, rdone, None, None, stream=st)
sim.prefetch_resets(stream=st); torch.cuda.synchronize()
words = np.zeros(1 << 16, np.uint64)
cat = {}
conv_all = []
slowest = []
for t in range(60):
    sim.step(acts[t], None, obs, rew, done, None, None, stream=st)
    torch.cuda.synchronize()
    ms = sim.last_kernel_ms()
    rc = sim._lib.srl_sim_get_state(sim.handle, 99, words.ctypes.data, words.nbytes)
    assert rc == 0
    nslots = ((n + 6) // 7) * 8
    w = words[:nslots]
    cyc = (w >> np.uint64(32)).astype(np.int64); fl = (w & np.uint64(0xFFFFFFFF)).astype(np.int64)
    live = cyc > 0
    conv_all.extend(((fl[live & ((fl & (1 << 18)) == 0)] >> 24) & 255).tolist())
    k = int(np.argmax(cyc))
    slowest.append((ms * 1e3, cyc[k], fl[k]))
    for c, f in zip(cyc[live], fl[live]):
        key = ("helper " if f & (1 << 18) else "") + ("done " if f & (1 << 16) else "") + ("record " if f & (1 << 17) else "") + \
              ("general(%d sweeps) " % ((f >> 8) & 255) if f & 2 else "") + ("watch " if f & 1 else "") + ("limit " if f & 64 else "") + "nc=%d" % ((f >> 2) & 15)
        a = cat.setdefault(key, [0, 0, 0]); a[0] += 1; a[1] += c; a[2] = max(a[2], c)
    words[:] = 0
print("slowest slot of each launch: launch us, slot cycles (us at 1.965 GHz), flags")
for ms, c, f in slowest[:30]:
    print("  %.1f us  %d cycles (%.1f us)  helper=%d done=%d record=%d general=%d(%d sweeps) watch=%d limit=%d nc=%d"
          % (ms, c, c / 1965.0, (f >> 18) & 1, (f >> 16) & 1, (f >> 17) & 1, (f >> 1) & 1, (f >> 8) & 255, f & 1, (f >> 6) & 1, (f >> 2) & 15))
conv = np.array(conv_all)
print("first sweep of the fast loop in which no motor row moved (0 = never within 150): histogram over %d slot-steps" % conv.size)
for lo, hi in ((0, 0), (1, 20), (21, 40), (41, 60), (61, 80), (81, 100), (101, 120), (121, 150)):
    print("  %3d..%3d: %6.2f %%" % (lo, hi, 100.0 * ((conv >= lo) & (conv <= hi)).mean()))
print("categories: count, mean exit time us, max us")
for key, (cnt, tot, mx) in sorted(cat.items(), key=lambda kv: -kv[1][1] / kv[1][0]):
    print("  %-60s %7d  %.1f  %.1f" % (key, cnt, tot / cnt / 1965.0, mx / 1965.0))
