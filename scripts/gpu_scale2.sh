#!/bin/bash
# 2-GPU check of the driver's launch line (one rank per GPU over NCCL), both arms and both workloads; run with `gpurun --gpus 2`
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale2_kuka.json
$TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload mobile 2>/dev/null | tail -1 > gpurun_out/scale2_mobile.json
$TR --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference 2>/dev/null | tail -1 > gpurun_out/scale2_reference.json
python -c "
import json
for w in ('kuka','mobile','reference'):
    d=json.load(open('gpurun_out/scale2_%s.json'%w)); print(w, d.get('impl','b200'), d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
