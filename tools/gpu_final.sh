# final GPU visit of the round: full parity suite (all packed plans present), smoke, bench (both arms), per-step table,
# ncu launch list of the bench command, DRAM traffic per conv launch, ncu --set full pages of the top kernels
python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_final.log 2>&1; tail -4 gpurun_out/pytest_gpu_final.log; grep -h "tf32: prob map\|max |prob" gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
VSE_STEP_TABLE=gpurun_out/steps_final.txt python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
VSE_PDL=0 python bench.py --no-cpu-baseline > gpurun_out/bench_final_nopdl.json 2>/dev/null; cut -c1-160 gpurun_out/bench_final_nopdl.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/b_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel --launch-skip 55 --launch-count 55 --csv --log-file gpurun_out/tc_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_traffic.log 2>&1
python tools/make_traffic_json.py gpurun_out/tc_traffic.csv gpurun_out/top_kernel_traffic.json | cut -c1-300
for sk in ${NCU_TC_SKIPS:-55 80}; do
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip $sk --launch-count 1 -o gpurun_out/tc_l$sk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_l$sk.log 2>&1
  ncu -i gpurun_out/tc_l$sk.ncu-rep --page raw --csv > gpurun_out/tc_final_l$sk.raw.csv 2>/dev/null
  ncu -i gpurun_out/tc_l$sk.ncu-rep --page details --csv > gpurun_out/tc_final_l$sk.details.csv 2>/dev/null
  rm -f gpurun_out/tc_l$sk.ncu-rep
done
for sk in ${NCU_DW_SKIPS:-28 34}; do
  ncu --set full --clock-control none --import-source on -k regex:dwconv_reg_kernel --launch-skip $sk --launch-count 1 -o gpurun_out/dw_l$sk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_dw_l$sk.log 2>&1
  ncu -i gpurun_out/dw_l$sk.ncu-rep --page raw --csv > gpurun_out/dw_l$sk.raw.csv 2>/dev/null
  rm -f gpurun_out/dw_l$sk.ncu-rep
done
ls gpurun_out | wc -l
