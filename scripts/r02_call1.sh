#!/bin/bash
# Round 2, GPU call 1: run the next-episode-record (reset prefetch) path for the first time on hardware:
# gated tests, compute-sanitizer memcheck on a side-stream lockstep run, lockstep launch timing, PPO2 phases.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
( SRL_TEST_PREFETCH=1 timeout 300 python -m pytest tests/test_prefetch_gpu.py -q -x 2>&1 | tail -25 ) > gpurun_out/c1_prefetch_pytest.txt
( timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_prefetch.py 2>&1 | tail -30 ) > gpurun_out/c1_memcheck.txt
( SRL_TEST_PREFETCH=1 timeout 120 python scripts/step_launch_timing.py 2>&1 | tail -10 ) > gpurun_out/c1_step_timing.txt
( timeout 300 python - <<'PY' 2>&1 | tail -8
from rl_baselines.ppo2 import train
n, T, updates = 4096, 128, 10
for kw in (dict(), dict(prefetch_resets=True), dict(prefetch_resets=True, fused_act=True), dict(fused_act=True)):
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
) > gpurun_out/c1_ppo2.txt
( timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > gpurun_out/c1_default_pytest.txt
( timeout 120 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/c1_bench.txt
tail -n 30 gpurun_out/c1_*.txt
