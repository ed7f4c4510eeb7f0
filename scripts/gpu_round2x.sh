#!/bin/bash
mkdir -p gpurun_out
PROBE_V3_ONLY=1 PROBE_ONLY=0 PROBE_SR=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc3 -s 2 -c 1 -o gpurun_out/prof_nerfmlp3 -f python scripts/probe_v3.py > gpurun_out/ncu_v3.log 2>&1; echo "ncu_v3 rc=$?"
PROBE_V3_ONLY=1 PROBE_SR=1 PROBE_TIMELINES=16 timeout 100 python scripts/probe_v3.py > gpurun_out/probe_v3.log 2>&1; cat gpurun_out/probe_v3.log | cut -c1-120
