#!/bin/bash
# usage: run_all_workloads.sh <tag>  -- every bench workload, the reference arm, launch list and ncu --set full of the two hot kernels
T=$1
python bench.py > gpurun_out/bench_default_$T.json 2> gpurun_out/bench_default_$T.err; echo default rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$T.json 2> gpurun_out/bench_reference_$T.err; echo reference rc=$?
for W in tiny longread filter bacterial; do
  python bench.py --workload $W --steps 10 --no-cpu-baseline > gpurun_out/bench_${W}_$T.json 2> gpurun_out/bench_${W}_$T.err; echo $W rc=$?
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gs_ --launch-skip 15 -c 40 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fastq > gpurun_out/ncu_l.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:"gs_label|gs_reduce_thread" --launch-skip 6 --launch-count 2 -o gpurun_out/prof_final_$T python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fastq --reads-per-step 1000000 > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?
python - <<PY
import json
for w in ("default","reference","tiny","longread","filter","bacterial"):
    try:
        d=json.load(open("gpurun_out/bench_%s_$T.json" % w))
        print(w, "value %.3g %s" % (d["value"], d["unit"]), "e2e %.3g" % d["e2e"]["value"], "frac", d.get("roofline",{}).get("frac"), "fq", (d.get("e2e_fastq") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(w, "ERR", e)
PY
