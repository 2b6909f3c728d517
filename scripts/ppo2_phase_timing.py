"""Where a PPO2 update of config 3 (KukaButtonGymEnv-v0, 4096 envs, n_steps 128) spends its time: collect / GAE / optimise (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
from rl_baselines.ppo2 import train
n, T, updates = 4096, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 10
for fused in (True,):
    pt = {}
    hist = train("KukaButtonGymEnv-v0", n, n * T * updates, seed=0, verbose=0, phase_times=pt, fused_act=fused)
    tot = sum(pt.values())
    print("fused_act=%s  phases over %d updates (synchronised): " % (fused, updates) + ", ".join("%s %.1f ms/update (%.0f %%)" % (k, 1e3 * v / updates, 100 * v / tot) for k, v in pt.items())
          + " -> %.2f M env-steps/s with the syncs" % (n * T * updates / tot / 1e6))
    hist = train("KukaButtonGymEnv-v0", n, n * T * updates, seed=0, verbose=0, fused_act=fused)
    print("fused_act=%s  unsynchronised: %.2f M env-steps/s cumulative incl. graph capture, mean return %.3f" % (fused, hist[-1][2] / 1e6, hist[-1][1]))
