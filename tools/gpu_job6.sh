for bo in 1 0; do
echo "== halo base offset mode $bo"
VSE_TC_HALO_BASEOFF=$bo python -m pytest tests -m gpu -x -q -k "plan_steps or fast_kernels or end_to_end_synthetic_fp16" 2>&1 | tail -4
done
VSE_STEP_TABLE=gpurun_out/steps.txt python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-300; tail -3 gpurun_out/bench.err
