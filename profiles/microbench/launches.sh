#!/bin/bash
# per-kernel durations of one timed step (ncu serialises: compare shares, not absolutes). usage: launches.sh <out.csv> "<bench flags>"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gs_ --launch-skip 15 --launch-count 5 --csv --log-file $1 python bench.py --no-cpu-baseline --steps 2 --warmup 3 $2 > gpurun_out/ncu_l.log 2>&1
python - $1 <<PY
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5 and r[0].isdigit()]
for r in rows: print("%-60s %10.3f ms"%(r[4][:60], float(r[-1])/1e6))
PY
