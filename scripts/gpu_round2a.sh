#!/bin/bash
# round 1, session 2, call A: full GPU parity suite + contract bench (adds the Mip-NeRF arm) + launch list incl. the training step
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
tail -n 12 gpurun_out/t_all.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print('headline %.1f Mrays/s (%s) e2e %.1f | chain %.1f fused %.1f | train %.1f | nerf %.2f M (%.0f TF) | mip %.2f M (%.0f TF)' % (d['value']/1e6, d['config']['path'], d['e2e']['value']/1e6, d['paths']['chain']['value']/1e6, d['paths']['fused']['value']/1e6, d['train']['value']/1e6, d['nerf']['value']/1e6, d['nerf']['roofline']['achieved'], d['mip']['value']/1e6, d['mip']['roofline']['achieved']))
PY
tail -3 gpurun_out/bench.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_all.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu_list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
