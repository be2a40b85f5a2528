# ragged KxK convolutions on the tensor-core path (server recognisers): parity cases + timing of V4/ch_rec and V2/ch_rec
timeout 600 python -m pytest tests -m gpu -x -q -k "other_shipped or bilstm or plan_steps" > gpurun_out/pytest_gpu_ragged.log 2>&1; tail -4 gpurun_out/pytest_gpu_ragged.log
for r in V4/ch_rec V2/ch_rec; do
  n=$(echo $r | tr '/' '_')
  VSE_STEP_TABLE=gpurun_out/steps_$n.txt python bench.py --rec $r --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_$n.json'))
print('$r', 'fps', round(b['value'],1), 'ms', round(b['ms_per_step'],3), 'stages', [round(x,3) for x in b['stage_ms_last_e2e_step']], 'e2e', round(b['e2e']['value'],1), b['roofline']['per_kernel_ms'])
PY
done
