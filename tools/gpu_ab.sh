# A/B in ONE box visit (boxes differ by several percent): bench with and without a set of VSE_FLAG_* switches
for fl in ${AB_FLAGS:-0 4096}; do
  VSE_STEP_TABLE=gpurun_out/steps_f$fl.txt python bench.py --no-cpu-baseline --flags $fl > gpurun_out/bench_f$fl.json 2>/dev/null
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_f$fl.json'))
print('flags $fl', 'fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1))
PY
done
