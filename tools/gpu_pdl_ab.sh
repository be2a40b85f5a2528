# programmatic dependent launch on/off in ONE box visit + the full parity suite with it on
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_pdl.log 2>&1; tail -5 gpurun_out/pytest_gpu_pdl.log
for m in ${PDL_MODES:-0 1}; do
  VSE_PDL=$m VSE_STEP_TABLE=gpurun_out/steps_pdl$m.txt timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl$m.json 2>gpurun_out/bench_pdl$m.err
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_pdl$m.json'))
print('pdl $m', 'fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1))
PY
done
