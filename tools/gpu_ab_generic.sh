# parity subset + bench of the current build (compare with the previous visit's numbers in profiles/)
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ab.log 2>&1; tail -3 gpurun_out/pytest_gpu_ab.log
VSE_STEP_TABLE=gpurun_out/steps_ab.txt python bench.py --no-cpu-baseline > gpurun_out/bench_ab.json 2>gpurun_out/bench_ab.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_ab.json'))
print('fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1))
print(b['roofline']['per_kernel_ms'])
PY
