#!/bin/bash
XRB_NERF_MLP_V=2 timeout 120 python -m pytest tests/test_gpu_nerf_mlp.py -q -m gpu -x 2>&1 | tail -n 5
timeout 200 python scripts/bench_nerf.py 2>&1 | tail -5
for d in 7 2 4; do echo "== dbg=$d (1 no-TMA, 2 no-MMA, 4 no-epilogue)"; XRB_NM_DBG=$d timeout 100 python scripts/bench_nerf.py 2>&1 | tail -2; done
