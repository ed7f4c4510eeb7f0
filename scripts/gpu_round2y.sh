#!/bin/bash
# final re-validation after the NerfMLP v3 mode work: full GPU suite, smoke, contract bench
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
tail -n 5 gpurun_out/t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
tail -n 2 gpurun_out/smoke.log
timeout 500 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print('headline %.1f Mrays/s (%s) e2e %.1f | chain %.1f fused %.1f | train %.1f | nerf %.2f M (%.0f TF) | mip %.2f M (%.0f TF)' % (d['value']/1e6, d['config']['path'], d['e2e']['value']/1e6, d['paths']['chain']['value']/1e6, d['paths']['fused']['value']/1e6, d['train']['value']/1e6, d['nerf']['value']/1e6, d['nerf']['roofline']['achieved'], d['mip']['value']/1e6, d['mip']['roofline']['achieved']))
print('parity', d['parity']); print('roofline', d['roofline'])
PY
tail -3 gpurun_out/bench.err
XRB_N3_SHARED_RING=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-train --no-grid > gpurun_out/bench_sr.json 2> gpurun_out/bench_sr.err; echo "bench_sr rc=$?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_sr.json').read().strip().split('\n')[-1])
print('private rings: nerf %.2f M (%.0f TF) | mip %.2f M (%.0f TF)' % (d['nerf']['value']/1e6, d['nerf']['roofline']['achieved'], d['mip']['value']/1e6, d['mip']['roofline']['achieved']))
PY
cat gpurun_out/summary.txt
