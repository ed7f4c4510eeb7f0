#!/bin/bash
mkdir -p gpurun_out
PROBE_V3_ONLY=1 PROBE_ONLY=0 timeout 120 python scripts/probe_v3.py > gpurun_out/probe_v3.log 2>&1; cat gpurun_out/probe_v3.log
