#!/bin/bash
# usage: instcount.sh <workload> : warp instructions and duration of the label kernel (1 M reads), default build and variants
for V in "" $GS_VARIANTS; do
  export GS_LIB_VARIANT=$V
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:gs_label --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/ic.csv python bench.py --workload $1 --steps 1 --warmup 3 --no-cpu-baseline --no-fastq --reads-per-step 1000000 > gpurun_out/ncu_t.log 2>&1
  echo "$1 variant=[$V]"; grep gs_label gpurun_out/ic.csv | awk -F'","' '{print "   ", $(NF-2), $NF}'
done
