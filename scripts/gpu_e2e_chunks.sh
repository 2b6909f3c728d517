#!/bin/bash
# srl_sim_rollout_host pipelining: parity tests, then e2e throughput against the number of T-chunks; ncu --set full of the MobileRobot kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_mobile_gpu.py tests/test_kuka_gpu.py -m gpu -q -x -k "rollout_host" 2>&1 | tail -3 | tee gpurun_out/e2e_pytest.txt
row() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"; }
for c in 1 0 4 16; do SRL_HOST_CHUNKS=$c python bench.py --workload mobile --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | row "mobile chunks=$c"; done | tee gpurun_out/e2e_chunks.txt
for c in 1 2 4 8; do SRL_HOST_CHUNKS=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | row "kuka chunks=$c"; done | tee -a gpurun_out/e2e_chunks.txt
ncu --set full --clock-control none --import-source on -k regex:mobile_rollout -s 4 -c 1 -o gpurun_out/r01b_mobile_full python bench.py --workload mobile --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
