#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
tail -n 4 gpurun_out/t_all.log
XRB_TC_REGS=128 timeout 300 python scripts/quick_bench.py 1
for cfg in "1 128" "2 128" "2 96" "3 96" "2 80" "4 96"; do
set -- $cfg
echo "== pipeline=$1 regs=$2"
XRB_TC_REGS=$2 timeout 300 python bench.py --steps 100 --warmup 10 --pipeline $1 2>gpurun_out/bench_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('value %.1f Mrays/s  ms/step %.3f  e2e %.1f Mrays/s  field_ms %.3f frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac']))
"
done
cat gpurun_out/summary.txt
