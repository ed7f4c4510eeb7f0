#!/bin/bash
mkdir -p gpurun_out
timeout 60 scripts/micro/lat > gpurun_out/lat.log 2>&1; cat gpurun_out/lat.log
timeout 60 scripts/micro/hop > gpurun_out/hop.log 2>&1; cat gpurun_out/hop.log
PROBE_V3_ONLY=1 timeout 120 python scripts/probe_v3.py > gpurun_out/probe_v3.log 2>&1; tail -4 gpurun_out/probe_v3.log
