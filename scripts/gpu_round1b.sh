#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu_list rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ngp_field_tc -s 3 -c 2 -o gpurun_out/prof_field -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu_full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 5 gpurun_out/t_all.log; tail -n 3 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench_ref.json
