#!/bin/bash
# round 1, session 2, call B: NerfMLP v3 (two tiles in flight) + specialised encoders: parity then speed
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 100 python -m pytest tests/test_gpu_nerf_mlp.py -q -m gpu -k v3 > gpurun_out/t_v3.log 2>&1; echo "pytest_v3 rc=$?" >> gpurun_out/summary.txt
tail -n 25 gpurun_out/t_v3.log
timeout 200 python -m pytest tests/test_gpu_nerf_mlp.py tests/test_gpu_nerf.py -q -m gpu -k 'not v3' > gpurun_out/t_nerf.log 2>&1; echo "pytest_nerf rc=$?" >> gpurun_out/summary.txt
tail -n 25 gpurun_out/t_nerf.log
XRB_SKIP_LIB=1 timeout 120 python scripts/bench_nerf.py > gpurun_out/bench_nerf.log 2>&1; echo "bench_nerf rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench_nerf.log | tail -8
for v in 2 3; do
XRB_NERF_MLP_V=$v timeout 150 python bench.py --steps 20 --warmup 3 --no-train > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "bench_v$v rc=$?" >> gpurun_out/summary.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v$v.json').read().strip().split('\n')[-1])
print('v$v: nerf %.2f M (%.0f TF, %.2f ms) | mip %.2f M (%.0f TF, %.2f ms)' % (d['nerf']['value']/1e6, d['nerf']['roofline']['achieved'], d['nerf']['ms_per_batch'], d['mip']['value']/1e6, d['mip']['roofline']['achieved'], d['mip']['ms_per_batch']))
PY
done
cat gpurun_out/summary.txt
