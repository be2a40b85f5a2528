timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "plan_steps_match or small_capacities or dense_frame or prefetch" > gpurun_out/pytest_split_a.log 2>&1; tail -5 gpurun_out/pytest_split_a.log
timeout 900 python -m pytest tests/test_gpu_real_video.py -m gpu -q -s > gpurun_out/pytest_gpu_real_video.log 2>&1; grep -E "^test_|passed|failed" gpurun_out/pytest_gpu_real_video.log | tail -14
for pr in fp32_tc fp16; do
VSE_STEP_TABLE=gpurun_out/steps_$pr.txt python bench.py --no-cpu-baseline --precision $pr > gpurun_out/bench_$pr.json 2>gpurun_out/bench_$pr.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_$pr.json'))
print('$pr fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1), 'lines/frame', b['text_lines_per_frame'], b['mean_padded_rec_width'])
print(b['roofline']['per_kernel_ms'])
PY
done
