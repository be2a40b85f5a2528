# one iteration on the GPU box: parity suite (what the snapshot holds), real-video + job tests, bench with the step table
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_parity.log 2>&1; tail -4 gpurun_out/pytest_parity.log
timeout 900 python -m pytest tests/test_gpu_real_video.py tests/test_gpu_jobs.py -m gpu -q -s > gpurun_out/pytest_video_jobs.log 2>&1; grep -E "^test_|accurate stretch|passed|failed|Error" gpurun_out/pytest_video_jobs.log | tail -16
VSE_STEP_TABLE=gpurun_out/steps_iter.txt python bench.py --no-cpu-baseline > gpurun_out/bench_iter.json 2>gpurun_out/bench_iter.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_iter.json'))
print('fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1), 'roofline', round(b['roofline']['frac'],3), b['roofline']['kernel'])
print(b['roofline']['per_kernel_ms'])
PY
