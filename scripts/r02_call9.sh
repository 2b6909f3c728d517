#!/bin/bash
# GPU call 9: full-batch parity statistics (configs 2, 5, action_repeat 3, force_down off), trainer logging tests, record tests, PPO2 phases
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 900 python -m pytest tests/test_kuka_gpu.py -q -s -k "full_batch or lane_packing" 2>&1 | grep -E "FULL-BATCH|passed|failed|Error|assert" | head -30 ) > gpurun_out/c9_full_batch.txt
( timeout 900 python -m pytest tests/test_trainer_gpu.py tests/test_prefetch_gpu.py tests/test_policy_gpu.py -q 2>&1 | tail -15 ) > gpurun_out/c9_trainer.txt
( timeout 400 python - <<'PY' 2>&1 | tail -8
from rl_baselines.ppo2 import train
n, T, updates = 4096, 128, 10
for kw in (dict(), dict(prefetch_resets=True), dict(prefetch_resets=True, fused_act=True)):
    try:
        pt = {}
        train("KukaButtonGymEnv-v0", n, n * T * updates, seed=0, verbose=0, phase_times=pt, **kw)
        tot = sum(pt.values())
        hist = train("KukaButtonGymEnv-v0", n, n * T * updates, seed=0, verbose=0, **kw)
        print(kw, ", ".join("%s %.1f ms" % (k, 1e3 * v / updates) for k, v in pt.items()), "-> %.2f M env-steps/s synchronised, %.2f M unsynchronised, return %.3f"
              % (n * T * updates / tot / 1e6, hist[-1][2] / 1e6, hist[-1][1]))
    except Exception as ex:
        print(kw, "FAILED", repr(ex))
PY
) > gpurun_out/c9_ppo2.txt
tail -n 30 gpurun_out/c9_*.txt | cut -c1-300
