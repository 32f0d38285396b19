#!/bin/bash
# usage: launches_wl.sh <workload> <tag> : per-kernel durations of one step of a bench workload
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gs_label|gs_reduce|gs_mark|gs_filter" --launch-skip 8 -c 12 --csv --log-file gpurun_out/launches_$1_$2.csv python bench.py --workload $1 --steps 2 --warmup 3 --no-cpu-baseline --no-fastq > gpurun_out/ncu_l.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_$1_$2.csv")) if len(r)>5]
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if r[0]=="ID": hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if d.get("Metric Name")=="gpu__time_duration.sum":
        v=float(d["Metric Value"].replace(",","")); u=d["Metric Unit"]
        v = v/1e6 if u.startswith("n") else v/1e3 if u.startswith("u") else v
        agg[d["Kernel Name"][:60]].append(v)
print("$1")
for k,v in sorted(agg.items(), key=lambda kv:-max(kv[1])): print("  %-62s n=%d min=%.3f max=%.3f ms" % (k,len(v),min(v),max(v)))
PY
