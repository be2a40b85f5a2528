# final GPU visit of the round: full parity suite (all packed plans present), smoke, bench (both arms), per-step table
python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_final.log 2>&1; tail -6 gpurun_out/pytest_gpu_final.log; grep -h "tf32: prob map\|max |prob" gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json | cut -c1-400
VSE_STEP_TABLE=gpurun_out/steps_final.txt python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
