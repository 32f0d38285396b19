#!/bin/bash
# (1) parity of the final flat filter kernel; (2) timing experiment: table addressed by minimizer (wrong results, label-kernel time only)
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_goals.py -x -q -k "filter" 2>&1 | tail -2
for v in fakelocal main; do
  LIBV=""; [ $v = fakelocal ] && LIBV=/root/repo/genestrip_b200/_lib/fakelocal.so
  GS_LIB_VARIANT=$LIBV timeout 900 python bench.py --steps 20 --warmup 5 --also longread --no-cpu-baseline --no-fastq > gpurun_out/r02/bench_n_$v.json 2> gpurun_out/r02/bench_n_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for v in ("fakelocal", "main"):
    try:
        j = json.load(open("gpurun_out/r02/bench_n_%s.json" % v))
        print(v, "viral value %.2f label %.3f ms reduce %.3f hits %d" % (j["value"] / 1e9, j["roofline"]["kernel_ms"], j["roofline"]["reduce_kernels_ms"], j["hits_total"]))
        for n, r in j["workloads"].items():
            print("  ", n, "value %.2f label %.3f ms reduce %.3f hits %d" % (r["value"] / 1e9, r["roofline"]["kernel_ms"], r["roofline"]["reduce_kernels_ms"], r["hits_total"]))
    except Exception as e:
        print(v, "ERR", e)
PY
