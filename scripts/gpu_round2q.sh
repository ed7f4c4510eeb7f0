#!/bin/bash
mkdir -p gpurun_out
PROBE_V3_ONLY=1 PROBE_ONLY=0 PROBE_TIMELINES=48,52,49,53 timeout 120 python scripts/probe_v3.py > gpurun_out/probe_v3.log 2>&1; grep -A 12 "timeline" gpurun_out/probe_v3.log | cut -c1-100
