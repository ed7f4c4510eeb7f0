#!/bin/bash
mkdir -p gpurun_out
for ls in 1 2 4; do
XRB_MARCH_LANE_STRIDE=$ls XRB_TC_REGS=128 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file gpurun_out/launches_ls$ls.csv python scripts/quick_bench.py 1 > gpurun_out/qb_ls$ls.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_ls$ls.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows: agg.setdefault(r[4].split('(')[0],[]).append(float(r[-1]))
print('LS=$ls', {k[-30:]:round(sum(v)/len(v)/1e3,1) for k,v in agg.items()})
PY
done
XRB_MARCH_LANE_STRIDE=1 timeout 300 python scripts/quick_bench.py 1
