mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; tail -c 300 gpurun_out/bench_ref_final.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 1 --no-ref-gpu > /dev/null 2>&1
python scripts/launch_list.py gpurun_out/launches_r02_final.csv "round 2 final: launch list of python bench.py --steps 2 --warmup 1 --no-ref-gpu (first 600 launches)" > gpurun_out/r02_final_launch_list.md 2>/dev/null; head -14 gpurun_out/r02_final_launch_list.md
XRB_FIELD_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:ngp_field_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_field_r02_final python scripts/field_bench.py 7 > /dev/null 2>&1; ls -la gpurun_out/prof_field_r02_final.ncu-rep
