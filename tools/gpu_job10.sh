ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel --launch-skip 55 --launch-count 55 --csv --log-file gpurun_out/tc_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/tc_traffic.csv | cut -c1-200
for sk in 55 80; do
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip $sk --launch-count 1 -o gpurun_out/tc_final_l$sk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 > gpurun_out/ncu_final_l$sk.log 2>&1
ncu -i gpurun_out/tc_final_l$sk.ncu-rep --page raw --csv > gpurun_out/tc_final_l$sk.raw.csv 2>/dev/null
ncu -i gpurun_out/tc_final_l$sk.ncu-rep --page details --csv > gpurun_out/tc_final_l$sk.details.csv 2>/dev/null
rm -f gpurun_out/tc_final_l$sk.ncu-rep
done
ls -la gpurun_out | tail -6
