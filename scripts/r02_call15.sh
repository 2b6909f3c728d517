#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 600 python -m pytest tests/test_render_gpu.py -q -s 2>&1 | grep -E "RENDER|passed|failed|Error|assert" | head -20 ) > gpurun_out/c15_render.txt
( timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_render_gpu.py -q -k "MobileRobot2Target or Kuka2Button" 2>&1 | tail -5 ) > gpurun_out/c15_memcheck.txt
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/c15_pytest.txt
( timeout 300 python bench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ncu_capture'], d['roofline']['frac'], d['roofline']['useful_lane_frac'], d['secondary']['mobile_config4']['value'], d['secondary']['mobile_config4']['roofline']['ncu_capture'])" ) > gpurun_out/c15_bench.txt 2>&1
cat gpurun_out/c15_render.txt gpurun_out/c15_memcheck.txt gpurun_out/c15_pytest.txt gpurun_out/c15_bench.txt
