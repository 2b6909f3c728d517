#!/bin/bash
# A/B timing of kernel code-shape variants (run on the GPU box)
cd "$(dirname "$0")/.."
for lib in robotics-rl-srl_b200/csrc/libsrl_variant_*.so; do
  echo "== $lib"
  SRL_SIM_CUDA_LIB=$PWD/$lib python scripts/kuka_prof.py 4096 128 4 7 2>&1 | tail -3
  SRL_SIM_CUDA_LIB=$PWD/$lib python scripts/kuka_prof.py 32768 128 3 32 2>&1 | tail -1
done
