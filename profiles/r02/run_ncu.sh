#!/bin/bash
# ncu evidence of round 2: launch lists and --set full captures at the bench's OWN launch size (4 M reads / 60 k long reads per
# launch), so that roofline.traffic needs no scaling.  Usage (on the GPU box): bash profiles/r02/run_ncu.sh [workloads...]
mkdir -p gpurun_out/r02
WLS=${@:-viral bacterial longread filter}
for W in $WLS; do
  K="gs_label"; [ $W = filter ] && K="gs_filter_flat"
  # plain run first (exit 0 without ncu), then the launch list, then the full capture of the dominant kernel
  python bench.py --workload $W --also none --steps 2 --warmup 3 --no-cpu-baseline --no-fastq > gpurun_out/r02/plain_$W.json 2> gpurun_out/r02/plain_$W.err || { echo "$W plain run failed"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gs_" --launch-skip 12 -c 24 --csv --log-file gpurun_out/r02/launches_$W.csv \
      python bench.py --workload $W --also none --steps 2 --warmup 3 --no-cpu-baseline --no-fastq > gpurun_out/r02/ncu_l_$W.log 2>&1; echo "$W launches rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 3 --launch-count 1 -o gpurun_out/r02/prof_$W -f \
      python bench.py --workload $W --also none --steps 1 --warmup 3 --no-cpu-baseline --no-fastq > gpurun_out/r02/ncu_f_$W.log 2>&1; echo "$W full rc=$?"
done
ls -la gpurun_out/r02/*.ncu-rep
