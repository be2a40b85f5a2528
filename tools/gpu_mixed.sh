# mixed mode (detector fp32 tensor-core mode, recogniser fp16): real-video parity + bench, beside the fp32_tc bench
timeout 600 python -m pytest tests/test_gpu_real_video.py -m gpu -q -s -k "mixed" > gpurun_out/pytest_mixed.log 2>&1; grep -E "^test_|passed|failed|Error" gpurun_out/pytest_mixed.log | tail
for m in fp32_tc mixed; do
VSE_STEP_TABLE=gpurun_out/steps_$m.txt python bench.py --no-cpu-baseline --precision $m > gpurun_out/bench_$m.json 2>gpurun_out/bench_$m.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_$m.json'))
print('$m fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1), 'roofline', round(b['roofline']['frac'],3), b['roofline']['kernel'], 'lines', b['text_lines_per_frame'])
print(b['roofline']['per_kernel_ms'])
PY
done
