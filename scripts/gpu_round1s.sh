#!/bin/bash
XRB_BENCH_DEBUG=1 timeout 300 python bench.py --steps 60 --warmup 10 --no-train --no-nerf --pipeline 1 2>&1 >/dev/null | grep "\[bench\]"
