#!/bin/bash
# round 1, session 2, final capture: full GPU parity suite, smoke, contract bench, launch list, ncu --set full of the NerfMLP v3 kernel
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
print('parity', d['parity']); print('grid', d['grid_update']); print('roofline', d['roofline'])
PY
tail -3 gpurun_out/bench.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> gpurun_out/summary.txt
tail -c 600 gpurun_out/bench_ref.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_all.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu_list rc=$?" >> gpurun_out/summary.txt
PROBE_V3_ONLY=1 PROBE_ONLY=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc3 -s 2 -c 1 -o gpurun_out/prof_nerfmlp3 -f python scripts/probe_v3.py > gpurun_out/ncu_v3.log 2>&1; echo "ncu_v3 rc=$?" >> gpurun_out/summary.txt
PROBE_V3_ONLY=1 timeout 100 python scripts/probe_v3.py > gpurun_out/probe_v3.log 2>&1; tail -16 gpurun_out/probe_v3.log
cat gpurun_out/summary.txt
