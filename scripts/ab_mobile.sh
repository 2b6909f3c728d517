#!/bin/bash
cd "$(dirname "$0")/.."
for lib in robotics-rl-srl_b200/csrc/libsrl_variant_pf*.so robotics-rl-srl_b200/csrc/libsrl_sim_b200.so; do
  echo "== $lib"
  SRL_SIM_CUDA_LIB=$PWD/$lib python bench.py --workload mobile --steps 20 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
