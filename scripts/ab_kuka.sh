#!/bin/bash
# A/B timing of Kuka kernel build variants on the default bench workload (run on the GPU box)
cd "$(dirname "$0")/.."
for lib in robotics-rl-srl_b200/csrc/libsrl_sim_b200.so robotics-rl-srl_b200/csrc/libsrl_variant_*.so; do
  SRL_SIM_CUDA_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $lib)', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
