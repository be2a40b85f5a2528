"""profiles/top_kernel_traffic.json from an ncu metrics pass over every conv_tc_kernel launch of one bench step:
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv_tc_kernel
    --launch-skip 55 --launch-count 55 --csv --log-file <csv> python bench.py --steps 1 --warmup 1 ..."""
import csv, json, sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr = None
per = {}
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = d["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(unit, 1)
        per.setdefault(int(d["ID"]), {})[d["Metric Name"]] = v * scale
n = len(per)
rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
t = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
out = {"kernel": "conv_tc_kernel", "launches": n, "dram_bytes_read_per_launch": rd / n, "dram_bytes_write_per_launch": wr / n,
       "dram_bytes_per_launch": (rd + wr) / n, "ncu_time_s_per_launch": t / n,
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over the "
              "55 conv_tc_kernel launches of one bench step (32 x 1080p frames, det + rec); cold-cache, serialised"}
json.dump(out, open(dst, "w"), indent=1)
print(out)
