for fl in 0 512; do
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 80 --launch-count 1 -o gpurun_out/tc3x3_f$fl -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 --flags $fl > gpurun_out/ncu_f$fl.log 2>&1
  ncu -i gpurun_out/tc3x3_f$fl.ncu-rep --page source --csv > gpurun_out/tc3x3_f$fl.source.csv 2>/dev/null
  ncu -i gpurun_out/tc3x3_f$fl.ncu-rep --page raw --csv > gpurun_out/tc3x3_f$fl.raw.csv 2>/dev/null
  rm -f gpurun_out/tc3x3_f$fl.ncu-rep
done
ls -la gpurun_out | tail -8
