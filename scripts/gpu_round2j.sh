#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
PROBE_V3_ONLY=1 PROBE_ONLY=0 timeout 120 python scripts/probe_v3.py > gpurun_out/probe_v3.log 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/probe_v3.log
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
tail -n 5 gpurun_out/t_all.log
timeout 500 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print('headline %.1f Mrays/s (%s) e2e %.1f | train %.1f | nerf %.2f M (%.0f TF) | mip %.2f M (%.0f TF)' % (d['value']/1e6, d['config']['path'], d['e2e']['value']/1e6, d['train']['value']/1e6, d['nerf']['value']/1e6, d['nerf']['roofline']['achieved'], d['mip']['value']/1e6, d['mip']['roofline']['achieved']))
print('parity', d['parity'])
PY
tail -3 gpurun_out/bench.err
cat gpurun_out/summary.txt
