#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/t_all.log
timeout 200 python scripts/bench_fused.py 2>&1 | tail -3
timeout 200 python scripts/quick_bench.py 1 2>&1 | tail -5
