#!/bin/bash
mkdir -p gpurun_out
for d in 1 2 4 3 6 7; do echo "== dbg=$d (1 no-TMA, 2 no-MMA, 4 no-epilogue)"; XRB_NM_DBG=$d timeout 100 python scripts/bench_nerf.py 2>&1 | tail -1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc2 -s 2 -c 1 -o gpurun_out/prof_nerfmlp2 -f env XRB_NM_DBG=0 python scripts/bench_nerf.py > gpurun_out/ncu_nm.log 2>&1; echo "ncu rc=$?"
