VSE_STEP_TABLE=gpurun_out/steps.txt python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-200; tail -3 gpurun_out/bench.err
VSE_DW_TH4=1 VSE_STEP_TABLE=gpurun_out/steps_th4.txt python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_th4.json 2>&1
echo th8; grep -h dwconv gpurun_out/steps.txt | awk '{print $3}' | tr '\n' ' '; echo; echo th4; grep -h dwconv gpurun_out/steps_th4.txt | awk '{print $3}' | tr '\n' ' '
