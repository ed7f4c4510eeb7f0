#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_render.py -q -x -k "fused" > gpurun_out/t_fused.log 2>&1; echo "pytest_fused rc=$?"
tail -n 5 gpurun_out/t_fused.log
XRB_DEBUG=1 timeout 200 python scripts/bench_fused.py 2>&1 | grep -v "occupancy query" | tail -6
for dbg in 1 2 3; do echo "== dbg $dbg"; XRB_FUSED_DBG=$dbg timeout 120 python scripts/bench_fused.py 65536 f 2>&1 | grep "P="; done
