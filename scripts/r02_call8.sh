#!/bin/bash
# GPU call 8: coop kernel with the tight sweep loop: variants, full parity suite, sanitizer, step timing, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( for lib in libsrl_sim_b200.so libsrl_variant_nodefer.so libsrl_variant_notight.so; do echo "$lib: $(SRL_SIM_CUDA_LIB=$PWD/robotics-rl-srl_b200/csrc/$lib timeout 60 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -1)"; done; echo "thread-per-env: $(SRL_KUKA_COOP=0 timeout 60 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -1)" ) > gpurun_out/c8_quick.txt
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/c8_pytest.txt
( timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python scripts/sanitize_run.py 2>&1 | tail -6 ) > gpurun_out/c8_racecheck.txt
( timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_prefetch.py 2>&1 | tail -6 ) > gpurun_out/c8_memcheck_prefetch.txt
( timeout 200 python scripts/step_launch_timing.py 2>&1 | tail -6 ) > gpurun_out/c8_step_timing.txt
( timeout 300 python bench.py --no-secondary 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'])" ) > gpurun_out/c8_bench.txt 2>&1
tail -n 32 gpurun_out/c8_*.txt | cut -c1-330
