#!/bin/bash
# GPU call: coop kernel quick A/B + ncu capture by function
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 60 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -1; SRL_SIM_CUDA_LIB=$PWD/robotics-rl-srl_b200/csrc/libsrl_variant_exit_late.so timeout 60 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -1; SRL_KUKA_COOP=0 timeout 60 python scripts/kuka_prof.py 4096 128 4 2>&1 | tail -1 ) > gpurun_out/c7_quick.txt
( timeout 600 python -m pytest tests/test_kuka_gpu.py -q 2>&1 | tail -4 ) > gpurun_out/c7_pytest.txt
KCMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-secondary"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:kuka_kernel -s 14 -c 1 -f -o gpurun_out/r02_kuka_coop $KCMD > gpurun_out/c7_ncu_kuka.log 2>&1 )
( timeout 120 python scripts/ncu_summary.py gpurun_out/r02_kuka_coop.ncu-rep ) > gpurun_out/c7_ncu_summary.txt 2>&1
( timeout 200 python scripts/ncu_by_line.py gpurun_out/r02_kuka_coop.ncu-rep 30 ) > gpurun_out/c7_ncu_by_function.txt 2>&1
tail -n 30 gpurun_out/c7_quick.txt gpurun_out/c7_pytest.txt | cut -c1-300; tail -n 8 gpurun_out/c7_ncu_summary.txt | cut -c1-250; head -45 gpurun_out/c7_ncu_by_function.txt | cut -c1-230
