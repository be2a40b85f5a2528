# fused depthwise -> pointwise: parity + bench with and without the fusion (vse_config.flags bit 16384 switches it off)
bash tools/gpu_iter.sh
VSE_STEP_TABLE=gpurun_out/steps_nofuse.txt python bench.py --no-cpu-baseline --flags 16384 > gpurun_out/bench_nofuse.json 2>gpurun_out/bench_nofuse.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_nofuse.json'))
print('NO FUSION fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'launches', b['gpu_launches'])
print(b['roofline']['per_kernel_ms'])
PY
