mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.json; echo
XRB_FIELD_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:ngp_field_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_field_r02_final python scripts/field_bench.py 7 > /dev/null 2>&1; ls -la gpurun_out/prof_field_r02_final.ncu-rep
