#!/bin/bash
# Round 2, second final GPU call (1 GPU), after the scaled general loop / incremental watch / tiled policy step / culled ray caster:
# full parity suite, sanitizer, ncu captures -> profiles JSON keyed by the instruction-stream hash, launch list, bench lines, timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "FULL-BATCH|passed|failed|Error|RENDER THROUGHPUT|finished episodes found" | tail -40 ) > gpurun_out/g_pytest.txt
KCMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-secondary"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:kuka_kernel -s 14 -c 1 -f -o gpurun_out/r02_kuka_full $KCMD > gpurun_out/g_ncu_kuka.log 2>&1 )
( timeout 120 python scripts/ncu_to_json.py kuka gpurun_out/r02_kuka_full.ncu-rep gpurun_out/r02_kuka_ncu.json "ncu --set full --clock-control none --import-source on -k regex:kuka_kernel -s 14 -c 1 $KCMD" ) > gpurun_out/g_ncu_kuka_json.txt 2>&1
( timeout 120 python scripts/ncu_summary.py gpurun_out/r02_kuka_full.ncu-rep ) > gpurun_out/r02_kuka_kernel_ncu_full.txt 2>&1
( timeout 200 python scripts/ncu_by_line.py gpurun_out/r02_kuka_full.ncu-rep 25 ) >> gpurun_out/r02_kuka_kernel_ncu_full.txt 2>&1
cp gpurun_out/r02_kuka_ncu.json profiles/ 2>/dev/null
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 )
( timeout 600 python bench.py 2>&1 | tail -1 ) > gpurun_out/r02_bench.json
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/r02_bench_reference.json
( timeout 300 python scripts/step_launch_timing.py 2>&1 | tail -10 ) > gpurun_out/g_step_timing.txt
( timeout 300 python scripts/kuka_prof.py 8192 128 3 2>&1 | tail -1; timeout 100 python scripts/kuka_prof.py 2048 128 3 2>&1 | tail -1; timeout 100 python scripts/kuka_prof.py 16384 128 3 2>&1 | tail -1; timeout 100 python scripts/kuka_prof.py 32768 128 3 2>&1 | tail -1 ) > gpurun_out/g_batch_scan.txt
( timeout 400 python scripts/ppo2_phase_timing.py 10 2>&1 | tail -4 ) > gpurun_out/g_ppo2.txt
( timeout 300 python scripts/render_timing.py 2>&1 ) > gpurun_out/g_render_timing.txt
( timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_render_gpu.py -x -q -k "match_the_cpu or culling" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" ) > gpurun_out/g_memcheck.txt
( timeout 600 compute-sanitizer --tool memcheck python scripts/consumer_kernels_for_ncu.py 2>&1 | grep -E "ERROR SUMMARY|done" ) >> gpurun_out/g_memcheck.txt
( timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_run.py 2>&1 | grep -E "ERROR SUMMARY|kuka dones|mobile ok" ) >> gpurun_out/g_memcheck.txt
cat gpurun_out/g_pytest.txt gpurun_out/g_step_timing.txt gpurun_out/g_batch_scan.txt gpurun_out/g_ppo2.txt gpurun_out/g_memcheck.txt; cut -c1-900 gpurun_out/r02_bench.json; head -24 gpurun_out/r02_kuka_kernel_ncu_full.txt | cut -c1-200
