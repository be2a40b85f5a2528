# ncu: launch list of one bench step, then --set full (+ source) captures of the longest launches of the kernels named in $NCU_KERNELS
# (regex fragments, space separated; default: conv_tc_kernel dwconv_reg_kernel).  Pages land in gpurun_out/ncu_<tag>_*.csv
TAG=${NCU_TAG:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 ${NCU_BENCH_ARGS:-}"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
python - <<PY
import csv, os, re
rows=[r for r in csv.reader(open('gpurun_out/launches_$TAG.csv')) if len(r)>5 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Metric Unit, Metric Value
names=[r[4] for r in rows]; dur=[float(r[-1].replace(',','')) for r in rows]
tot=sum(dur)
agg={}
for n,d in zip(names,dur):
    k=re.sub(r'\(.*','',n); k=re.sub(r'<.*','',k); agg[k]=agg.get(k,0)+d
print('launches',len(rows),'total us',round(tot/1e3,1))
for k,v in sorted(agg.items(),key=lambda kv:-kv[1])[:14]: print(f'  {k:40s} {v/1e3:9.1f} us {100*v/tot:5.1f}%')
plan=[]
for frag in os.environ.get('NCU_KERNELS','conv_tc_kernel dwconv_reg_kernel').split():
    idx=[i for i,n in enumerate(names) if frag in n]
    half=[i for i in idx if i>=len(rows)//2] or idx       # the timed step is the second half of the list
    best=sorted(half,key=lambda i:-dur[i])[:int(os.environ.get('NCU_TOP','2'))]
    for i in best:
        skip=sum(1 for j in idx if j<i)
        plan.append((frag,skip,round(dur[i]/1e3,1)))
open('gpurun_out/ncu_plan.txt','w').write('\n'.join(f'{a} {b} {c}' for a,b,c in plan)+'\n')
print(plan)
PY
while read frag skip us; do
  ncu --set full --clock-control none --import-source on -k regex:$frag --launch-skip $skip --launch-count 1 -o gpurun_out/cap_${frag}_$skip -f $CMD > gpurun_out/ncu_cap.log 2>&1
  ncu -i gpurun_out/cap_${frag}_$skip.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_${frag}_${skip}.raw.csv 2>/dev/null
  ncu -i gpurun_out/cap_${frag}_$skip.ncu-rep --page details --csv > gpurun_out/ncu_${TAG}_${frag}_${skip}.details.csv 2>/dev/null
  ncu -i gpurun_out/cap_${frag}_$skip.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_${frag}_${skip}.source.csv 2>/dev/null
  rm -f gpurun_out/cap_${frag}_$skip.ncu-rep
  echo "captured $frag skip $skip ($us us)"
done < gpurun_out/ncu_plan.txt
