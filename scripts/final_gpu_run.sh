#!/bin/bash
# Round-end evidence run (one B200): GPU test suite, both bench workloads, e2e A/B of the zero-copy host outputs, divergence report
# (configs 2 / 5), ncu launch lists of the bench commands, smoke.  Outputs land in gpurun_out/ (copied into profiles/ afterwards).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/final_pytest.txt
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_bench_kuka.json
SRL_HOST_ZEROCOPY=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_bench_kuka_staged.json
python bench.py --workload mobile --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_bench_mobile.json
python -c "
import json
for w in ('kuka','kuka_staged','mobile'):
    d=json.load(open('gpurun_out/final_bench_%s.json'%w)); print(w, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline'].get('issue',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'), d['clocks'])" | tee gpurun_out/final_bench_summary.txt
timeout 300 python scripts/divergence_report.py > /dev/null 2>gpurun_out/divergence_err.txt; tail -3 gpurun_out/divergence_report.txt
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches_kuka.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches_mobile.csv python bench.py --workload mobile --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
