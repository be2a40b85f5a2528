# A/B of the depthwise kernel families in ONE box visit (boxes differ by several percent) + the parity cases they touch
python -m pytest tests -m gpu -x -q -k "other_shipped or fast_kernels or plan_steps" > gpurun_out/pytest_gpu_dw.log 2>&1; tail -5 gpurun_out/pytest_gpu_dw.log
for m in ${DW_MODES:-1 2}; do
  VSE_DW_MODE=$m VSE_STEP_TABLE=gpurun_out/steps_dw$m.txt python bench.py --no-cpu-baseline > gpurun_out/bench_dw$m.json 2>gpurun_out/bench_dw$m.err
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_dw$m.json'))
print('dw_mode $m', 'fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1))
PY
done
