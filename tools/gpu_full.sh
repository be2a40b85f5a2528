# the driver's end-of-round view: every GPU test (all packed plans + videos present), smoke, default bench + reference arm
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_full.log 2>&1; grep -E "^test_|accurate stretch|passed|failed|Error|skipped" gpurun_out/pytest_gpu_full.log | tail -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --config 2 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; VSE_STEP_TABLE=gpurun_out/steps_cfg2.txt python bench.py --config 2 --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import json
b=json.load(open('gpurun_out/bench_cfg2.json'))
print('cfg2 fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1), 'roofline', b['roofline']['bound'], round(b['roofline']['frac'],3), round(b['roofline']['achieved'],1), b['roofline']['unit'])
print(b['roofline']['per_kernel_ms'])
PY
