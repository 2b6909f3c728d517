#!/bin/bash
# First GPU call of the next round (one B200, ~3 min): validate the experimental next-episode records (written at the end of round 1 without a GPU),
# measure what they do to the lockstep step and to the PPO2 update, and re-check the default suite.  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( SRL_TEST_PREFETCH=1 timeout 200 python -m pytest tests/test_prefetch_gpu.py -q -x 2>&1 | tail -15 ) | tee gpurun_out/prefetch_pytest.txt
SRL_TEST_PREFETCH=1 timeout 90 python scripts/step_launch_timing.py 2>&1 | tail -8 | tee gpurun_out/prefetch_step_timing.txt
timeout 120 python - <<'PY' 2>&1 | tail -6 | tee gpurun_out/prefetch_ppo2.txt
from rl_baselines.ppo2 import train
n, T, updates = 4096, 128, 10
for kw in (dict(), dict(prefetch_resets=True), dict(prefetch_resets=True, fused_act=True)):
    pt = {}
    train("KukaButtonGymEnv-v0", n, n * T * updates, seed=0, verbose=0, phase_times=pt, **kw)
    tot = sum(pt.values())
    hist = train("KukaButtonGymEnv-v0", n, n * T * updates, seed=0, verbose=0, **kw)
    print(kw, ", ".join("%s %.1f ms" % (k, 1e3 * v / updates) for k, v in pt.items()), "-> %.2f M env-steps/s synchronised, %.2f M unsynchronised, return %.3f"
          % (n * T * updates / tot / 1e6, hist[-1][2] / 1e6, hist[-1][1]))
PY
( timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) | tee gpurun_out/default_pytest.txt
