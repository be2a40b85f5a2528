# source-level ncu pages (stall samples per SASS line) of conv_tc_kernel launches: skip index = launch number inside one bench step
for sk in ${NCU_TC_SKIPS:-55 80}; do
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip $sk --launch-count 1 -o gpurun_out/tc_l$sk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_l$sk.log 2>&1
  ncu -i gpurun_out/tc_l$sk.ncu-rep --page source --csv > gpurun_out/tc_src_l$sk.source.csv 2>/dev/null
  rm -f gpurun_out/tc_l$sk.ncu-rep
done
ls -la gpurun_out/tc_src_*
