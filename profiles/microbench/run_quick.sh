#!/bin/bash
# usage: run_quick.sh <tag>  -- gpu tests, bench, launch list
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_$1.json 2> gpurun_out/v.err
python -c "
import json; d=json.load(open('gpurun_out/v_$1.json')); print('value %.2f G  e2e %.2f G  kernel_ms %.3f' % (d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['kernel_ms']))"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gs_ --launch-skip 15 -c 12 --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_$1.csv")) if len(r)>5]
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if r[0]=="ID": hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if d.get("Metric Name")=="gpu__time_duration.sum":
        v=float(d["Metric Value"].replace(",","")); u=d["Metric Unit"]
        v = v/1e6 if u.startswith("n") else v/1e3 if u.startswith("u") else v
        agg[d["Kernel Name"][:60]].append(v)
for k,v in sorted(agg.items(), key=lambda kv:-max(kv[1])): print("%-62s n=%d min=%.3f max=%.3f ms" % (k,len(v),min(v),max(v)))
PY
