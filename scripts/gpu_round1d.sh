#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python scripts/quick_bench.py 1 > gpurun_out/qb_tc.log 2>&1; echo "qb rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_count -s 2 -c 1 -o gpurun_out/prof_march -f python scripts/quick_bench.py 1 > gpurun_out/ncu_march.log 2>&1; echo "ncu_march rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 8 gpurun_out/t_all.log; cat gpurun_out/qb_tc.log
