#!/bin/bash
# closing verification of the committed state: full GPU suite + contract bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?"; tail -n 3 gpurun_out/t_all.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print('headline %.1f e2e %.1f train %.1f' % (d['value']/1e6, d['e2e']['value']/1e6, d['train']['value']/1e6))
for k in ('nerf','nerf_train','mip'): print(k, d[k].get('value'), d[k].get('error'))
print('image fused', d['image']['fused']['value'], d['image']['fused']['roofline']['frac']); print('parity', d['parity']['chain'])
PY
