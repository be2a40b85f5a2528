python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -8 gpurun_out/pytest_gpu_all.log
python bench.py --no-cpu-baseline --det V4/ch_det --rec V4/ch_rec --flags 1024 --steps 3 --warmup 1 > gpurun_out/bench_server.json 2> gpurun_out/bench_server.err; cat gpurun_out/bench_server.json | cut -c1-1200; tail -3 gpurun_out/bench_server.err
