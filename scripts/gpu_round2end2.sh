#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 3 --no-train --no-grid --no-image --no-mip > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().split('\n')[-1])
print('headline %.1f' % (d['value']/1e6)); print('nerf', d['nerf'].get('value'), d['nerf'].get('cpu_baseline')); print('nerf_train', d['nerf_train'])
PY
