#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --warmup 3 --no-train --no-grid --no-image --no-mip --no-nerf > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().split('\n')[-1])
print('headline %.1f' % (d['value']/1e6)); print('chain', d['paths']['chain']['value']/1e6, 'fused', d['paths']['fused']['value']/1e6, 'e2e', d['e2e']['value']/1e6)
PY
