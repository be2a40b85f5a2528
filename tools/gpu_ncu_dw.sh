# ncu --set full of two depthwise launches of one bench step: launch index by -k regex + skip (28 dw launches per step)
for sk in ${NCU_SKIPS:-34}; do
  ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-dwconv_reg_kernel} --launch-skip $sk --launch-count 1 -o gpurun_out/dw_l$sk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_dw_l$sk.log 2>&1
  ncu -i gpurun_out/dw_l$sk.ncu-rep --page raw --csv > gpurun_out/dw_l$sk.raw.csv 2>/dev/null
  ncu -i gpurun_out/dw_l$sk.ncu-rep --page source --csv > gpurun_out/dw_l$sk.source.csv 2>/dev/null
  rm -f gpurun_out/dw_l$sk.ncu-rep
done
ls gpurun_out | grep dw_
