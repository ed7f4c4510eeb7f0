#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --no-train --no-nerf --no-mip --no-grid > gpurun_out/bench_img.json 2> gpurun_out/bench_img.err; echo "rc=$?"
tail -3 gpurun_out/bench_img.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_img.json').read().strip().split('\n')[-1])
print('headline %.1f' % (d['value']/1e6)); print('image', d['image'])
PY
