# round-2 first visit: full GPU parity suite (server plans + real videos present), bench with step table
timeout 900 python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_real_video.py > gpurun_out/pytest_gpu_r2a.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2a.log
timeout 600 python -m pytest tests/test_gpu_real_video.py -m gpu -q -s > gpurun_out/pytest_gpu_real_video.log 2>&1; tail -15 gpurun_out/pytest_gpu_real_video.log
VSE_STEP_TABLE=gpurun_out/steps_r2a.txt python bench.py --no-cpu-baseline > gpurun_out/bench_r2a.json 2>gpurun_out/bench_r2a.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_r2a.json'))
print('fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1))
print(b['roofline']['per_kernel_ms'])
PY
