#!/bin/bash
# Round 2, GPU call 5: first run of the four-lanes-per-env kernel (kuka_coop.cuh): sanitizer, parity suite, A/B against one thread per env
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_run.py 2>&1 | tail -14 ) > gpurun_out/c5_memcheck.txt
( timeout 400 compute-sanitizer --tool racecheck --print-limit 10 python scripts/sanitize_run.py 2>&1 | tail -14 ) > gpurun_out/c5_racecheck.txt
( timeout 60 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -2; SRL_KUKA_COOP=0 timeout 60 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -1 ) > gpurun_out/c5_quick.txt
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/c5_pytest.txt
( for co in 1 0; do echo "SRL_KUKA_COOP=$co: $(SRL_KUKA_COOP=$co timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])")"; done ) > gpurun_out/c5_ab.txt 2>&1
( timeout 200 python scripts/step_launch_timing.py 2>&1 | tail -6 ) > gpurun_out/c5_step_timing.txt
tail -n 45 gpurun_out/c5_*.txt | cut -c1-400
