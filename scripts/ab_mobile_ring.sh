#!/bin/bash
# A/B of the MobileRobot action staging variants (register ring vs cp.async shared-memory ring of depth 3/4/6); run on the GPU box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests/test_mobile_gpu.py -m gpu -q -x 2>&1 | tail -3 ) 2>&1 | tee gpurun_out/ab_mobile_pytest.txt
for lib in robotics-rl-srl_b200/csrc/libsrl_variant_regring.so robotics-rl-srl_b200/csrc/libsrl_variant_pf_ring3.so robotics-rl-srl_b200/csrc/libsrl_sim_b200.so robotics-rl-srl_b200/csrc/libsrl_variant_pf_ring6.so; do
  for rep in 1 2; do
    SRL_SIM_CUDA_LIB=$PWD/$lib python bench.py --workload mobile --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $lib)', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
  done
done | tee gpurun_out/ab_mobile_ring.txt
