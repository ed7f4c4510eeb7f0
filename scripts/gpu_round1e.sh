#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_raymarch.py tests/test_gpu_render.py tests/test_gpu_field.py -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
XRB_TC_REGS=128 timeout 300 python scripts/quick_bench.py 1 > gpurun_out/qb_128.log 2>&1; echo "qb128 rc=$?" >> gpurun_out/summary.txt
XRB_TC_REGS=80 timeout 300 python scripts/quick_bench.py 1 > gpurun_out/qb_80.log 2>&1; echo "qb80 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 4 gpurun_out/t_all.log; cat gpurun_out/qb_128.log gpurun_out/qb_80.log
