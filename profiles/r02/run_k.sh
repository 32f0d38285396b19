#!/bin/bash
# label-kernel instruction diet: parity, then A/B of the library before (base.so) and after on the three match workloads
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_match.py -x -q 2>&1 | tail -3
for v in new base; do
  LIBV=""; [ $v = base ] && LIBV=/root/repo/genestrip_b200/_lib/base.so
  GS_LIB_VARIANT=$LIBV timeout 900 python bench.py --steps 20 --warmup 5 --also longread,bacterial --no-cpu-baseline --no-fastq > gpurun_out/r02/bench_k_$v.json 2> gpurun_out/r02/bench_k_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for v in ("new", "base"):
    try:
        j = json.load(open("gpurun_out/r02/bench_k_%s.json" % v))
        print(v, "viral value %.2f e2e %.2f label %.3f ms reduce %.3f" % (j["value"] / 1e9, j["e2e"]["value"] / 1e9, j["roofline"]["kernel_ms"], j["roofline"]["reduce_kernels_ms"]), j["hits_total"], j["unique_kmers_total"])
        for n, r in j["workloads"].items():
            print("  ", n, "value %.2f e2e %.2f label %.3f ms reduce %.3f" % (r["value"] / 1e9, r["e2e"]["value"] / 1e9, r["roofline"]["kernel_ms"], r["roofline"]["reduce_kernels_ms"]), r["hits_total"], r["unique_kmers_total"])
    except Exception as e:
        print(v, "ERR", e)
PY
