python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
VSE_STEP_TABLE=gpurun_out/steps.txt python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
VSE_DW_MODE=0 VSE_STEP_TABLE=gpurun_out/steps_dw0.txt python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_dw0.json 2>&1
echo tiled; grep -h dwconv gpurun_out/steps.txt | awk '{print $3}' | tr '\n' ' '; echo; echo strip; grep -h dwconv gpurun_out/steps_dw0.txt | awk '{print $3}' | tr '\n' ' '
