#!/bin/bash
# Round 2, GPU call 2: FMA issue micro-benchmark, A/B of the saturating sweep variants, parity suite on the new default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD/robotics-rl-srl_b200:$PYTHONPATH"
( timeout 120 scripts/microbench/fma_issue 2>&1 ) > gpurun_out/c2_fma_issue.txt
( timeout 600 bash scripts/ab_kuka.sh 2>&1 ) > gpurun_out/c2_ab.txt
( timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/c2_pytest.txt
for epw in 4 5 6 7 8; do echo "epw $epw: $(timeout 100 python scripts/kuka_prof.py 4096 128 4 $epw 2>&1 | tail -1)"; done > gpurun_out/c2_epw.txt
tail -n 40 gpurun_out/c2_*.txt
