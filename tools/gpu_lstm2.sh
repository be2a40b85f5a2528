timeout 300 python -m pytest tests -m gpu -x -q -k "bilstm or (other_shipped and V2)" 2>&1 | tail -4
VSE_STEP_TABLE=gpurun_out/steps_V2_ch_rec.txt timeout 300 python bench.py --rec V2/ch_rec --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_V2_ch_rec.json 2> gpurun_out/bench_V2_ch_rec.err
grep -i lstm gpurun_out/steps_V2_ch_rec.txt; cut -c1-200 gpurun_out/bench_V2_ch_rec.json
