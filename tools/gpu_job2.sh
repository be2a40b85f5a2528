python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
VSE_STEP_TABLE=gpurun_out/steps.txt python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for pr in 0 1 3; do
VSE_TC_L2PROMO=$pr VSE_STEP_TABLE=gpurun_out/steps_promo$pr.txt python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_promo$pr.json 2>&1
done
grep -h conv_tc gpurun_out/steps.txt | head -3; for pr in 0 1 3; do grep -h conv_tc gpurun_out/steps_promo$pr.txt | head -3; done
