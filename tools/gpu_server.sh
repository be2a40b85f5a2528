# server models (V4/ch_det + V4/ch_rec): per-step parity, accurate-mode job, config 2 bench with its step table
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "other_shipped or server or plan_steps" > gpurun_out/pytest_server.log 2>&1; tail -3 gpurun_out/pytest_server.log
timeout 900 python -m pytest tests/test_gpu_jobs.py -m gpu -q -s -k accurate > gpurun_out/pytest_acc.log 2>&1; grep -E "accurate stretch|passed|failed|Error" gpurun_out/pytest_acc.log | tail -5
python bench.py --config 2 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; VSE_STEP_TABLE=gpurun_out/steps_cfg2.txt python bench.py --config 2 --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import json
b=json.load(open('gpurun_out/bench_cfg2.json'))
print('cfg2 fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1), 'roofline', b['roofline']['bound'], round(b['roofline']['frac'],3), round(b['roofline']['achieved'],1), b['roofline']['unit'])
print(b['roofline']['per_kernel_ms'])
PY
