#!/bin/bash
# Round 2, final GPU call (1 GPU): full parity suite, sanitizer, ncu captures -> profiles JSON keyed by SASS hash, launch lists, bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "FULL-BATCH|passed|failed|Error|finished episodes found" | tail -30 ) > gpurun_out/f_pytest.txt
KCMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-secondary"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:kuka_kernel -s 14 -c 1 -f -o gpurun_out/r02_kuka_full $KCMD > gpurun_out/f_ncu_kuka.log 2>&1 )
( timeout 120 python scripts/ncu_to_json.py kuka gpurun_out/r02_kuka_full.ncu-rep gpurun_out/r02_kuka_ncu.json "ncu --set full --clock-control none --import-source on -k regex:kuka_kernel -s 14 -c 1 $KCMD" ) > gpurun_out/f_ncu_kuka_json.txt 2>&1
( timeout 120 python scripts/ncu_summary.py gpurun_out/r02_kuka_full.ncu-rep ) > gpurun_out/r02_kuka_kernel_ncu_full.txt 2>&1
( timeout 200 python scripts/ncu_by_line.py gpurun_out/r02_kuka_full.ncu-rep 25 ) >> gpurun_out/r02_kuka_kernel_ncu_full.txt 2>&1
MCMD="python bench.py --workload mobile --steps 12 --warmup 3 --no-cpu-baseline"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:mobile_rollout_kernel -s 14 -c 1 -f -o gpurun_out/r02_mobile_full $MCMD > gpurun_out/f_ncu_mobile.log 2>&1 )
( timeout 120 python scripts/ncu_to_json.py mobile gpurun_out/r02_mobile_full.ncu-rep gpurun_out/r02_mobile_ncu.json "ncu --set full --clock-control none --import-source on -k regex:mobile_rollout_kernel -s 14 -c 1 $MCMD" ) > gpurun_out/f_ncu_mobile_json.txt 2>&1
( timeout 120 python scripts/ncu_summary.py gpurun_out/r02_mobile_full.ncu-rep ) > gpurun_out/r02_mobile_rollout_ncu_full.txt 2>&1
mkdir -p profiles && cp gpurun_out/r02_kuka_ncu.json gpurun_out/r02_mobile_ncu.json profiles/ 2>/dev/null
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 )
( timeout 600 python bench.py 2>&1 | tail -1 ) > gpurun_out/r02_bench.json
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/r02_bench_reference.json
( timeout 200 python scripts/step_launch_timing.py 2>&1 | tail -6 ) > gpurun_out/f_step_timing.txt
( timeout 300 python scripts/kuka_prof.py 8192 128 3 2>&1 | tail -1; timeout 100 python scripts/kuka_prof.py 2048 128 3 2>&1 | tail -1; timeout 100 python scripts/kuka_prof.py 16384 128 3 2>&1 | tail -1 ) > gpurun_out/f_batch_scan.txt
cat gpurun_out/f_pytest.txt gpurun_out/f_step_timing.txt gpurun_out/f_batch_scan.txt; cut -c1-700 gpurun_out/r02_bench.json; head -24 gpurun_out/r02_kuka_kernel_ncu_full.txt | cut -c1-250
