#!/bin/bash
# round 2, call 46: GAE kernel + final full validation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|Error" | tail -12 ) > gpurun_out/c46_pytest.txt
( timeout 400 python scripts/ppo2_phase_timing.py 12 2>&1 | tail -3 ) > gpurun_out/c46_ppo2.txt
( timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 ) > gpurun_out/c46_smoke.txt
( timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330 ) > gpurun_out/c46_bench.txt
cat gpurun_out/c46_pytest.txt gpurun_out/c46_ppo2.txt gpurun_out/c46_smoke.txt gpurun_out/c46_bench.txt
