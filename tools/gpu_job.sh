# one GPU-box visit: parity tests, bench (+ per-step table), single-launch ncu captures of the tcgen05 conv kernel
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
VSE_STEP_TABLE=gpurun_out/steps.txt python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for sk in ${NCU_SKIPS:-55 76 80}; do
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip $sk --launch-count 1 -o gpurun_out/tc_l$sk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_l$sk.log 2>&1
  ncu -i gpurun_out/tc_l$sk.ncu-rep --page source --csv > gpurun_out/tc_l$sk.source.csv 2>/dev/null
  ncu -i gpurun_out/tc_l$sk.ncu-rep --page raw --csv > gpurun_out/tc_l$sk.raw.csv 2>/dev/null
done
ls -la gpurun_out
