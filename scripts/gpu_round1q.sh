#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
tail -n 6 gpurun_out/t_all.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('render %.1f Mrays/s e2e %.1f  train %.1f Mrays/s  nerf %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['train']['value']/1e6, d['nerf']))"
tail -3 gpurun_out/bench.err
cat gpurun_out/summary.txt
