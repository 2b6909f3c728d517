#!/bin/bash
# round 2, call 20: two pixels per thread, precomputed ray coefficients -- parity, timing, ncu of the raster kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_render_gpu.py -x -q -s > gpurun_out/c20_render.txt 2>&1; echo "render rc=$?" >> gpurun_out/c20_render.txt
timeout 300 python scripts/render_timing.py > gpurun_out/c20_render_timing.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -c 2 -o gpurun_out/c20_raster python scripts/render_timing.py > gpurun_out/c20_ncu.log 2>&1
tail -5 gpurun_out/c20_render.txt; cat gpurun_out/c20_render_timing.txt
