python -m pytest tests -m gpu -q -k "plan_steps or other_shipped or overflow" 2>&1 | tail -12
python bench.py --no-cpu-baseline --det V4/ch_det --rec V4/en_rec_fast --flags 2048 --steps 3 --warmup 1 > gpurun_out/bench_server_tf32.json 2> gpurun_out/bench_server_tf32.err; cat gpurun_out/bench_server_tf32.json | cut -c1-1100; tail -3 gpurun_out/bench_server_tf32.err
