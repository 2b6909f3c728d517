#!/bin/bash
# Round 2, GPU call 4: the new bench line (secondary legs, computed roofline), ncu captures -> profiles JSON keyed by SASS hash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 300 python -m pytest tests/test_prefetch_gpu.py -q -x 2>&1 | tail -4 ) > gpurun_out/c4_prefetch_pytest.txt
KCMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-secondary"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:kuka_kernel -s 14 -c 1 -f -o gpurun_out/r02_kuka_full $KCMD > gpurun_out/c4_ncu_kuka.log 2>&1 )
( timeout 120 python scripts/ncu_to_json.py kuka gpurun_out/r02_kuka_full.ncu-rep gpurun_out/r02_kuka_ncu.json "ncu --set full --clock-control none --import-source on -k regex:kuka_kernel -s 14 -c 1 $KCMD" ) > gpurun_out/c4_ncu_kuka_json.txt 2>&1
( timeout 120 python scripts/ncu_summary.py gpurun_out/r02_kuka_full.ncu-rep ) > gpurun_out/r02_kuka_kernel_ncu_full.txt 2>&1
MCMD="python bench.py --workload mobile --steps 12 --warmup 3 --no-cpu-baseline"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:mobile_rollout_kernel -s 14 -c 1 -f -o gpurun_out/r02_mobile_full $MCMD > gpurun_out/c4_ncu_mobile.log 2>&1 )
( timeout 120 python scripts/ncu_to_json.py mobile gpurun_out/r02_mobile_full.ncu-rep gpurun_out/r02_mobile_ncu.json "ncu --set full --clock-control none --import-source on -k regex:mobile_rollout_kernel -s 14 -c 1 $MCMD" ) > gpurun_out/c4_ncu_mobile_json.txt 2>&1
mkdir -p profiles && cp gpurun_out/r02_kuka_ncu.json gpurun_out/r02_mobile_ncu.json profiles/ 2>/dev/null
( timeout 600 python bench.py 2>&1 | tail -2 ) > gpurun_out/c4_bench.txt
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 ) > gpurun_out/c4_bench_ref.txt
tail -n 12 gpurun_out/c4_*.txt | cut -c1-3000
