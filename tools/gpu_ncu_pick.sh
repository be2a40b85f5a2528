# ncu --set full (+ source) of chosen launches of one bench step.  NCU_PICK = "tag:kernel-name-regex:ordinal" entries (space
# separated): the ordinal-th launch (0-based, from the start of the process) of the kernels matching the regex is captured — e.g. in
# V4/ch_det_fast conv_tc_kernel #25 is plan step 79 (3x3 96->24 at 136x240), #6 step 14 (1x1 192->192 at 34x60); dwconv_reg_kernel
# #6 is step 13 (5x5, 192 channels).  Pages land in gpurun_out/ncu_<NCU_TAG>_<tag>.{raw,details,source}.csv
TAG=${NCU_TAG:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pool 1 ${NCU_BENCH_ARGS:-}"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
python - <<PY
import csv, os, re
rows=[r for r in csv.reader(open('gpurun_out/launches_$TAG.csv')) if len(r)>5 and r[0].isdigit()]
names=[r[4] for r in rows]; dur=[float(r[-1].replace(',',''))/1e3 for r in rows]
tot=sum(dur); agg={}
for n,d in zip(names,dur):
    k=re.sub(r'\(.*','',n); k=re.sub(r'<.*','',k); agg[k]=agg.get(k,0)+d
print('launches',len(rows),'total us',round(tot,1))
for k,v in sorted(agg.items(),key=lambda kv:-kv[1])[:12]: print(f'  {k:40s} {v:9.1f} us {100*v/tot:5.1f}%')
plan=[]
for ent in os.environ.get('NCU_PICK','').split():
    tag,sub,ordn=ent.split(':'); ordn=int(ordn)
    idx=[i for i,n in enumerate(names) if re.search(sub,n)]
    if len(idx)<=ordn: print('no launch matches',ent); continue
    best=idx[ordn]
    plan.append((tag,sub,ordn,round(dur[best],1)))
    print('pick',tag,names[best][:110],'ordinal',ordn,'us',round(dur[best],1))
open('gpurun_out/ncu_plan.txt','w').write('\n'.join(f'{a} {b} {c} {d}' for a,b,c,d in plan)+'\n')
PY
while read tag kname skip us; do
  [ -z "$tag" ] && continue
  ncu --set full --clock-control none --import-source on -k regex:$kname --launch-skip $skip --launch-count 1 -o gpurun_out/cap_$tag -f $CMD > gpurun_out/ncu_cap.log 2>&1
  ncu -i gpurun_out/cap_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_$tag.raw.csv 2>/dev/null
  ncu -i gpurun_out/cap_$tag.ncu-rep --page details --csv > gpurun_out/ncu_${TAG}_$tag.details.csv 2>/dev/null
  case $tag in *_src) ncu -i gpurun_out/cap_$tag.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_$tag.source.csv 2>/dev/null;; esac   # (16 MB per conv page)
  rm -f gpurun_out/cap_$tag.ncu-rep
  echo "captured $tag ($kname skip $skip, $us us)"
done < gpurun_out/ncu_plan.txt
