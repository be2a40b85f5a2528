# V2/ch_rec (ResNet-34 + BiLSTM) behind the bench pipeline: per-step table with the recurrent kernel's time, and an
# ncu --set full page of one lstm_recurrent_kernel launch
VSE_STEP_TABLE=gpurun_out/steps_v2rec.txt python bench.py --rec V2/ch_rec --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_v2rec.json 2> gpurun_out/bench_v2rec.err; cut -c1-700 gpurun_out/bench_v2rec.json; tail -2 gpurun_out/bench_v2rec.err
grep -i "lstm" gpurun_out/steps_v2rec.txt
ncu --set full --clock-control none --import-source on -k regex:lstm_recurrent_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/lstm -f python bench.py --rec V2/ch_rec --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_lstm.log 2>&1
ncu -i gpurun_out/lstm.ncu-rep --page raw --csv > gpurun_out/lstm.raw.csv 2>/dev/null
ncu -i gpurun_out/lstm.ncu-rep --page source --csv > gpurun_out/lstm.source.csv 2>/dev/null
rm -f gpurun_out/lstm.ncu-rep; ls -la gpurun_out/lstm.*
