#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_registry.py -q -m gpu > gpurun_out/t_reg.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/t_reg.log
