# server models (V4/ch_det + V4/ch_rec): parity per step and a timing of BASELINE configs[2]-like work; needs the big plans
python -m pytest tests -m gpu -x -q -k "other_shipped" > gpurun_out/pytest_other.log 2>&1; tail -6 gpurun_out/pytest_other.log
VSE_STEP_TABLE=gpurun_out/steps_server.txt python bench.py --no-cpu-baseline --det V4/ch_det --rec V4/ch_rec --steps 5 --warmup 2 > gpurun_out/bench_server.json 2> gpurun_out/bench_server.err; cat gpurun_out/bench_server.json | cut -c1-1500; tail -3 gpurun_out/bench_server.err
