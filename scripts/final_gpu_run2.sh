#!/bin/bash
# Last GPU call of the round: whole GPU suite (incl. the fused policy-step kernels), PPO2 phase timing (torch path vs fused_act), both benches, smoke.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/final2_pytest.txt
timeout 70 python scripts/ppo2_phase_timing.py 10 2>&1 | tail -5 | tee gpurun_out/final2_ppo2_phases.txt
timeout 60 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final2_bench_kuka.json
timeout 40 python bench.py --workload mobile --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final2_bench_mobile.json
python -c "
import json
for w in ('kuka','mobile'):
    d=json.load(open('gpurun_out/final2_bench_%s.json'%w)); print(w, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline'].get('issue',{}).get('frac'))"
timeout 40 python __graft_entry__.py smoke 2>&1 | tail -1
