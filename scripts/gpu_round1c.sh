#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
XRB_DEBUG=1 timeout 300 python scripts/quick_bench.py 1 > gpurun_out/qb_tc.log 2>&1; echo "qb rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 15 gpurun_out/t_all.log; cat gpurun_out/qb_tc.log; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
