#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print('headline %.1f e2e %.1f' % (d['value']/1e6, d['e2e']['value']/1e6)); print('nerf', d['nerf']); print('nerf_train', d['nerf_train']); print('mip', str(d['mip'])[:200]); print('grid', str(d['grid_update'])[:150]); print('image', str(d['image'])[:200])
PY
