#!/bin/bash
# L2-persisting access window over the minimizer prefilter: A/B (GS_L2_PERSIST=0 switches it off)
mkdir -p gpurun_out/r02
for v in on off; do
  P=1; [ $v = off ] && P=0
  GS_L2_PERSIST=$P timeout 900 python bench.py --steps 20 --warmup 5 --also longread,bacterial --no-cpu-baseline --no-fastq > gpurun_out/r02/bench_l_$v.json 2> gpurun_out/r02/bench_l_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for v in ("on", "off"):
    try:
        j = json.load(open("gpurun_out/r02/bench_l_%s.json" % v))
        print(v, "viral value %.2f e2e %.2f label %.3f ms reduce %.3f" % (j["value"] / 1e9, j["e2e"]["value"] / 1e9, j["roofline"]["kernel_ms"], j["roofline"]["reduce_kernels_ms"]), j["native_options"].get("l2_persisting_window"))
        for n, r in j["workloads"].items():
            print("  ", n, "value %.2f e2e %.2f label %.3f ms reduce %.3f" % (r["value"] / 1e9, r["e2e"]["value"] / 1e9, r["roofline"]["kernel_ms"], r["roofline"]["reduce_kernels_ms"]), r["native_options"].get("l2_persisting_window"))
    except Exception as e:
        print(v, "ERR", e)
PY
