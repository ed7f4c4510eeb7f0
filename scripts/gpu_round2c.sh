#!/bin/bash
# round 1, session 2, call C: attribution of the NerfMLP v3 time + one ncu --set full capture of it; contract bench with parity / grid-update arms
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 150 python scripts/probe_v3.py > gpurun_out/probe_v3.log 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/probe_v3.log
PROBE_ONLY=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc3 -s 2 -c 1 -o gpurun_out/prof_nerfmlp3 -f python scripts/probe_v3.py > gpurun_out/ncu_v3.log 2>&1; echo "ncu_v3 rc=$?" >> gpurun_out/summary.txt
XRB_NERF_MLP_V=3 timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print('headline %.1f Mrays/s (%s) e2e %.1f | train %.1f | nerf %.2f M | mip %.2f M' % (d['value']/1e6, d['config']['path'], d['e2e']['value']/1e6, d['train']['value']/1e6, d['nerf']['value']/1e6, d['mip']['value']/1e6))
print('parity', d['parity']); print('grid', d['grid_update'])
PY
tail -5 gpurun_out/bench.err
cat gpurun_out/summary.txt
