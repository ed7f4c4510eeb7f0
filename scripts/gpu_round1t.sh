#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
tail -n 3 gpurun_out/t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print('headline %.1f Mrays/s (%s) e2e %.1f | chain %.1f fused %.1f | train %.1f | nerf %.2f M (%.0f TF)' % (d['value']/1e6, d['config']['path'], d['e2e']['value']/1e6, d['paths']['chain']['value']/1e6, d['paths']['fused']['value']/1e6, d['train']['value']/1e6, d['nerf']['value']/1e6, d['nerf']['roofline']['achieved']))
print('roofline', d['roofline'])
PY
tail -3 gpurun_out/bench.err
# launch list of the contract bench (both paths inside)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-train --no-nerf > gpurun_out/ncu_bench.log 2>&1; echo "ncu_list rc=$?" >> gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ngp_render_fused -s 3 -c 1 -o gpurun_out/prof_fused -f python bench.py --steps 4 --warmup 3 --no-train --no-nerf --pipeline 1 > gpurun_out/ncu_fused.log 2>&1; echo "ncu_fused rc=$?" >> gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ngp_field_tc -s 3 -c 1 -o gpurun_out/prof_field -f python bench.py --steps 4 --warmup 3 --no-train --no-nerf --pipeline 1 > gpurun_out/ncu_field.log 2>&1; echo "ncu_field rc=$?" >> gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:march_count -s 3 -c 1 -o gpurun_out/prof_march -f python bench.py --steps 4 --warmup 3 --no-train --no-nerf --pipeline 1 > gpurun_out/ncu_march.log 2>&1; echo "ncu_march rc=$?" >> gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc2 -s 2 -c 1 -o gpurun_out/prof_nerfmlp2 -f python scripts/bench_nerf.py > gpurun_out/ncu_nm.log 2>&1; echo "ncu_nerfmlp rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
