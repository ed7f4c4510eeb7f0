#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 400 python -m pytest tests/test_gpu_render.py -q -x -k "fused" > gpurun_out/t_fused.log 2>&1; echo "pytest_fused rc=$?" >> gpurun_out/summary.txt
tail -n 25 gpurun_out/t_fused.log
timeout 200 python scripts/bench_fused.py > gpurun_out/bench_fused.log 2>&1; echo "bench_fused rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench_fused.log | tail -8
timeout 300 python -m pytest tests/test_gpu_nerf.py -q -x -k "ray_generation" > gpurun_out/t_raygen.log 2>&1; echo "pytest_raygen rc=$?" >> gpurun_out/summary.txt
tail -n 5 gpurun_out/t_raygen.log
cat gpurun_out/summary.txt
