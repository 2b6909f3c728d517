#!/bin/bash
# A/B timing of MobileRobot kernel variants (chunk length x CTA size); run on the GPU box
cd "$(dirname "$0")/.."
for lib in robotics-rl-srl_b200/csrc/libsrl_variant_pf*.so robotics-rl-srl_b200/csrc/libsrl_sim_b200.so; do
  for blk in 0 32 64 128; do
    echo "== $lib block=$blk"
    SRL_MOBILE_BLOCK=$blk SRL_SIM_CUDA_LIB=$PWD/$lib python bench.py --workload mobile --steps 20 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
  done
done
