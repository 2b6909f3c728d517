#!/bin/bash
# Round 2, GPU call 6: four-lanes-per-env kernel after the ILP / linear-addressing pass: quick A/B + one ncu capture with source
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 60 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -1; SRL_KUKA_COOP=0 timeout 60 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -1 ) > gpurun_out/c6_quick.txt
( timeout 600 python -m pytest tests/test_kuka_gpu.py -q -x 2>&1 | tail -8 ) > gpurun_out/c6_pytest.txt
KCMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-secondary"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:kuka_kernel -s 14 -c 1 -f -o gpurun_out/r02_kuka_coop $KCMD > gpurun_out/c6_ncu_kuka.log 2>&1 )
( timeout 120 python scripts/ncu_summary.py gpurun_out/r02_kuka_coop.ncu-rep ) > gpurun_out/c6_ncu_summary.txt 2>&1
tail -n 30 gpurun_out/c6_quick.txt gpurun_out/c6_pytest.txt gpurun_out/c6_ncu_summary.txt | cut -c1-300
