#!/bin/bash
# A/B timing of Kuka kernel variants on the default bench workload (run on the GPU box)
cd "$(dirname "$0")/.."
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'])"; }
SRL_SIM_CUDA_LIB=$PWD/robotics-rl-srl_b200/csrc/libsrl_variant_base.so run base
SRL_KUKA_CTASYNC=0 run new_sync0
SRL_KUKA_CTASYNC=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('new_sync1', d['value'], d['ms_per_step'], d['e2e']['value'])"
