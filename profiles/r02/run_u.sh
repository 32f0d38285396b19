#!/bin/bash
mkdir -p gpurun_out/r02
for v in 1 0 1b; do
  OV=1; [ $v = 0 ] && OV=0
  env GS_OVERLAP_REDUCE=$OV timeout 900 python bench.py --steps 50 --warmup 5 --also longread,bacterial --no-cpu-baseline --no-fastq > gpurun_out/r02/bench_u_$v.json 2> gpurun_out/r02/bench_u_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for v in ("1", "0", "1b"):
    try:
        j = json.load(open("gpurun_out/r02/bench_u_%s.json" % v))
        print("overlap", v, "viral value %.2f e2e %.2f label %.3f reduce %.3f step %.3f" % (j["value"] / 1e9, j["e2e"]["value"] / 1e9, j["roofline"]["kernel_ms"], j["roofline"]["reduce_kernels_ms"], j["ms_per_step"]), j["hits_total"], j["unique_kmers_total"])
        for n, r in j["workloads"].items():
            print("    ", n, "value %.2f e2e %.2f label %.3f reduce %.3f step %.3f" % (r["value"] / 1e9, r["e2e"]["value"] / 1e9, r["roofline"]["kernel_ms"], r["roofline"]["reduce_kernels_ms"], r["ms_per_step"]), r["hits_total"], r["unique_kmers_total"])
    except Exception as e:
        print(v, "ERR", e)
PY
