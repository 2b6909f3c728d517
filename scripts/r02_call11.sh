#!/bin/bash
# GPU call 11 (2 GPUs): the bench line under torchrun (secondary legs across ranks), smoke, reference arm under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/c11_smoke.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -2 ) > gpurun_out/c11_bench2.txt
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 ) > gpurun_out/c11_bench2_ref.txt
cat gpurun_out/c11_smoke.txt; cut -c1-1500 gpurun_out/c11_bench2.txt; cut -c1-600 gpurun_out/c11_bench2_ref.txt
