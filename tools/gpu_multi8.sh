# 8-GPU box: default bench at N = 1 (same-box reference) and N = 8 with the balanced frame sets, then BASELINE configs[4]
# (4K synthetic 50k-frame stream, accurate models, frame-parallel)
python bench.py --no-cpu-baseline > gpurun_out/bench_n1_box8.json 2> gpurun_out/bench_n1_box8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
python - <<PY
import json
a=json.load(open('gpurun_out/bench_n1_box8.json')); b=json.load(open('gpurun_out/bench_n8.json'))
print('N=1 fps', round(a['value'],1), 'e2e', round(a['e2e']['value'],1))
print('N=8 fps', round(b['value'],1), 'e2e', round(b['e2e']['value'],1), 'eff', round(b['value']/8/a['value'],3), 'e2e eff', round(b['e2e']['value']/8/a['e2e']['value'],3), 'ms', round(b['ms_per_step'],3), 'per-rank dev ms', [r['device_ms_per_step'] for r in b['per_rank']], 'lines', [r['text_lines_per_step'] for r in b['per_rank']], 'h2d', round(b['e2e']['h2d_gbs'],1))
PY
if [ -n "$WITH_CFG4" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29509 bench.py --gpus 8 --config 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n8.json 2> gpurun_out/bench_cfg4_n8.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_cfg4_n8.json'))
print('cfg4 N=8 fps', round(b['value'],1), 'e2e', round(b['e2e']['value'],1), 'steps', b['steps'], 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'h2d', round(b['e2e']['h2d_gbs'],1))
PY
tail -2 gpurun_out/bench_cfg4_n8.err
fi
