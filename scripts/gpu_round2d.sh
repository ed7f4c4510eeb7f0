#!/bin/bash
# round 1, session 2, call D: NerfMLP v3 with hardware-barrier hand-offs (6 polling threads per SM instead of 20): parity, attribution, arms
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 100 python -m pytest tests/test_gpu_nerf_mlp.py -q -m gpu -k v3 > gpurun_out/t_v3.log 2>&1; rc=$?; echo "pytest_v3 rc=$rc" >> gpurun_out/summary.txt
tail -n 8 gpurun_out/t_v3.log
if [ $rc -eq 0 ]; then
PROBE_V3_ONLY=1 timeout 120 python scripts/probe_v3.py > gpurun_out/probe_v3.log 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/probe_v3.log
XRB_NERF_MLP_V=3 timeout 200 python bench.py --steps 20 --warmup 3 --no-train --no-grid > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo "bench_v3 rc=$?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v3.json').read().strip().split('\n')[-1])
print('v3: nerf %.2f M (%.0f TF, %.2f ms) | mip %.2f M (%.0f TF, %.2f ms)' % (d['nerf']['value']/1e6, d['nerf']['roofline']['achieved'], d['nerf']['ms_per_batch'], d['mip']['value']/1e6, d['mip']['roofline']['achieved'], d['mip']['ms_per_batch']))
print('parity', d['parity'])
PY
PROBE_ONLY=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc3 -s 2 -c 1 -o gpurun_out/prof_nerfmlp3b -f python scripts/probe_v3.py > gpurun_out/ncu_v3.log 2>&1; echo "ncu_v3 rc=$?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
