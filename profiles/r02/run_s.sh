#!/bin/bash
# label kernel compiled for 6 / 7 / 8 resident CTAs per SM after the instruction diet (40 / 36 / 32 registers)
mkdir -p gpurun_out/r02
for v in lb6 lb7 main; do
  LIBV=""; [ $v != main ] && LIBV=/root/repo/genestrip_b200/_lib/$v.so
  GS_LIB_VARIANT=$LIBV timeout 900 python bench.py --steps 30 --warmup 5 --also longread,bacterial --no-cpu-baseline --no-fastq > gpurun_out/r02/bench_s_$v.json 2> gpurun_out/r02/bench_s_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for v in ("lb6", "lb7", "main"):
    try:
        j = json.load(open("gpurun_out/r02/bench_s_%s.json" % v))
        print(v, "viral value %.2f label %.3f ms" % (j["value"] / 1e9, j["roofline"]["kernel_ms"]), end=" | ")
        for n, r in j["workloads"].items():
            print(n, "value %.2f label %.3f ms" % (r["value"] / 1e9, r["roofline"]["kernel_ms"]), end=" | ")
        print()
    except Exception as e:
        print(v, "ERR", e)
PY
