#!/bin/bash
XRB_NERF_MLP_V=2 timeout 120 python -m pytest tests/test_gpu_nerf_mlp.py -q -m gpu -x 2>&1 | tail -n 5
timeout 200 python scripts/bench_nerf.py 2>&1 | head -3
