# 4-GPU box: scaling of the default bench at N = 1, 2, 4 and BASELINE configs[3] (four videos, frame-range sharded)
for n in 1 2 4; do
  if [ $n = 1 ]; then python bench.py --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; fi
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_n$n.json'))
print('N=$n fps', round(b['value'],1), 'e2e', round(b['e2e']['value'],1), 'ms', round(b['ms_per_step'],3), 'dev', round(b['device_ms_per_step'],3), 'per-rank ms', [r['ms_per_step'] for r in b['per_rank']], 'h2d', round(b['e2e']['h2d_gbs'],1))
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29507 bench.py --gpus 4 --config 3 > gpurun_out/bench_cfg3_n4.json 2> gpurun_out/bench_cfg3_n4.err; cat gpurun_out/bench_cfg3_n4.json | cut -c1-1500; tail -3 gpurun_out/bench_cfg3_n4.err
python bench.py --config 3 > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err; cat gpurun_out/bench_cfg3_n1.json | cut -c1-900
