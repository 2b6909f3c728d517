#!/bin/bash
# Round 2, GPU call 3: helper-CTA next-episode records (tests, memcheck, lockstep timing, PPO2 phases), sweep variants A/B, full suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 400 python -m pytest tests/test_prefetch_gpu.py -q -x -s 2>&1 | tail -25 ) > gpurun_out/c3_prefetch_pytest.txt
( timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_prefetch.py 2>&1 | tail -12 ) > gpurun_out/c3_memcheck.txt
( timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_prefetch.py 2>&1 | tail -12 ) > gpurun_out/c3_racecheck.txt
( timeout 200 python scripts/step_launch_timing.py 2>&1 | tail -10 ) > gpurun_out/c3_step_timing.txt
( timeout 600 bash scripts/ab_kuka.sh 2>&1 ) > gpurun_out/c3_ab.txt
( timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/c3_pytest.txt
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
) > gpurun_out/c3_ppo2.txt
tail -n 40 gpurun_out/c3_*.txt
