# final GPU visit of round 2 (what the driver runs, plus the evidence for profiles/): every GPU test with all plans and videos,
# smoke, both bench arms, per-step table, ncu launch list, DRAM traffic per conv launch, two ncu --set full pages
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_final.log 2>&1; grep -E "^test_|accurate stretch|passed|failed|skipped|Error" gpurun_out/pytest_gpu_final.log | tail -16
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-200 gpurun_out/bench_reference.json; echo
VSE_STEP_TABLE=gpurun_out/steps_final.txt python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json; echo; tail -3 gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/b_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel --launch-skip 55 --launch-count 55 --csv --log-file gpurun_out/tc_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_traffic.log 2>&1
python tools/make_traffic_json.py gpurun_out/tc_traffic.csv gpurun_out/top_kernel_traffic.json | cut -c1-300
NCU_TAG=r02final NCU_PICK="conv3x3:conv_tc_kernel:25 dw5x5:dwconv_reg_kernel:6" bash tools/gpu_ncu_pick.sh 2>&1 | tail -4
