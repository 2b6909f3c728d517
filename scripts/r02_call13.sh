#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( for n in 4096 8192 16384 32768; do echo "$n envs: $(timeout 200 python scripts/kuka_prof.py $n 128 4 2>&1 | tail -1)"; done; echo "4096 envs, one thread per env: $(SRL_KUKA_COOP=0 timeout 100 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -1)" ) > gpurun_out/c13_batch_scan.txt
( timeout 900 python -m pytest tests/test_trainer_gpu.py tests/test_policy_gpu.py -q 2>&1 | tail -5; timeout 600 python -m pytest tests/test_kuka_gpu.py -q -k "no_force_down" 2>&1 | tail -3 ) > gpurun_out/c13_pytest.txt
( timeout 400 python bench.py 2>&1 | tail -1 ) > gpurun_out/c13_bench.json
( timeout 400 python - <<'PY' 2>&1 | tail -6
from rl_baselines.ppo2 import train
n, T, updates = 4096, 128, 12
for kw in (dict(fused_act=False, prefetch_resets=False), dict(fused_act=False), dict()):
    pt = {}
    train("KukaButtonGymEnv-v0", n, n * T * updates, seed=0, verbose=0, phase_times=pt, **kw)
    tot = sum(pt.values())
    best = 0
    for rep in range(2):
        hist = train("KukaButtonGymEnv-v0", n, n * T * updates, seed=0, verbose=0, **kw); best = max(best, hist[-1][2])
    print(kw or "defaults (records + fused policy step)", ", ".join("%s %.1f ms" % (k, 1e3 * v / updates) for k, v in pt.items()), "-> %.2f M env-steps/s synchronised, %.2f M unsynchronised (best of 2), return %.3f"
          % (n * T * updates / tot / 1e6, best / 1e6, hist[-1][1]))
PY
) > gpurun_out/c13_ppo2.txt
cat gpurun_out/c13_batch_scan.txt gpurun_out/c13_pytest.txt gpurun_out/c13_ppo2.txt; cut -c1-300 gpurun_out/c13_bench.json
