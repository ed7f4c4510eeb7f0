#!/bin/bash
# first GPU contact: parity of raymarch ops + field (CUDA-core first, tcgen05 under its own timeout), then timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_raymarch.py -q -m gpu -x > gpurun_out/t_raymarch.log 2>&1; echo "raymarch rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_field.py -q -m gpu -k "layout or modules or (vs_oracle and 0)" > gpurun_out/t_field_simt.log 2>&1; echo "field_simt rc=$?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_field.py -q -m gpu -k "(vs_oracle and 1) or strided or deeper" > gpurun_out/t_field_tc.log 2>&1; echo "field_tc rc=$?" >> gpurun_out/summary.txt
timeout 300 python scripts/quick_bench.py 0 > gpurun_out/qb_simt.log 2>&1; echo "qb_simt rc=$?" >> gpurun_out/summary.txt
timeout 300 python scripts/quick_bench.py 1 > gpurun_out/qb_tc.log 2>&1; echo "qb_tc rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_raymarch.log gpurun_out/t_field_simt.log gpurun_out/t_field_tc.log gpurun_out/qb_simt.log gpurun_out/qb_tc.log
