#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 90 python -m pytest tests/test_gpu_nerf_mlp.py -q -m gpu -k "cta_pair" -x > gpurun_out/t_v3.log 2>&1; rc=$?; echo "pytest_v3 rc=$rc" >> gpurun_out/summary.txt
tail -n 12 gpurun_out/t_v3.log
if [ $rc -eq 0 ]; then
PROBE_V3_ONLY=1 PROBE_ONLY=0 PROBE_SR=1 PROBE_TIMELINES=16 timeout 120 python scripts/probe_v3.py > gpurun_out/probe_v3.log 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/probe_v3.log | cut -c1-110
fi
cat gpurun_out/summary.txt
