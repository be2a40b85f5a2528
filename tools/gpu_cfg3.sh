# config 3 on one GPU, cold pass (warm-up 0) and warm pass (warm-up 1), with the phase breakdown; then the default bench
for w in 0 1; do
python bench.py --config 3 --warmup $w --steps 1 > gpurun_out/bench_cfg3_n1_w$w.json 2> gpurun_out/bench_cfg3_n1_w$w.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_cfg3_n1_w$w.json'))
print('cfg3 warmup $w fps', round(b['value'],1), 's/step', round(b['ms_per_step']/1e3,2), b['per_rank'][0], b['limiter'])
PY
done
VSE_STEP_TABLE=gpurun_out/steps_iter.txt python bench.py --no-cpu-baseline > gpurun_out/bench_iter.json 2>gpurun_out/bench_iter.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_iter.json'))
print('fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'roofline', round(b['roofline']['frac'],3))
print(b['roofline']['per_kernel_ms'])
PY
