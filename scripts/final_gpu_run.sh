#!/bin/bash
# Round-end evidence run (one B200): GPU test suite, both bench workloads, ncu launch lists of the bench commands, ncu --set full of the
# MobileRobot kernel, smoke.  Outputs land in gpurun_out/ (copied into profiles/ afterwards).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/final_pytest.txt
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_bench_kuka.json
python bench.py --workload mobile --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_bench_mobile.json
python -c "
import json
for w in ('kuka','mobile'):
    d=json.load(open('gpurun_out/final_bench_%s.json'%w)); print(w, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline'].get('issue',{}).get('frac'), 'cpu', d['cpu_baseline']['value'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches_kuka.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches_mobile.csv python bench.py --workload mobile --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mobile_rollout -s 4 -c 1 -o gpurun_out/final_mobile_full python bench.py --workload mobile --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
