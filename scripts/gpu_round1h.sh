#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_nerf_mlp.py -q -m gpu -x > gpurun_out/t_nerfmlp.log 2>&1; echo "nerf_mlp rc=$?" >> gpurun_out/summary.txt
tail -n 15 gpurun_out/t_nerfmlp.log
timeout 300 python scripts/bench_nerf.py > gpurun_out/bench_nerf.log 2>&1; echo "bench_nerf rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench_nerf.log | tail -5
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_nerf_mlp.py > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
tail -n 6 gpurun_out/t_all.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('render %.1f Mrays/s e2e %.1f  train %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['train']))"
tail -3 gpurun_out/bench.err
cat gpurun_out/summary.txt
