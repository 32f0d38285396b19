#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in a b; do
  timeout 900 python bench.py --steps 50 --warmup 5 --also longread,bacterial --no-cpu-baseline --no-fastq > gpurun_out/r02/bench_w_$v.json 2> gpurun_out/r02/bench_w_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for v in ("a", "b"):
    try:
        j = json.load(open("gpurun_out/r02/bench_w_%s.json" % v))
        hp = j["e2e"]["host_pack"]
        print(v, "viral value %.2f e2e %.2f (share %.2f, host %.2f ms) label %.3f reduce %.3f step %.3f" % (j["value"] / 1e9, j["e2e"]["value"] / 1e9, hp["packed_share_of_batch"], hp["host_ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["reduce_kernels_ms"], j["ms_per_step"]))
        for n, r in j["workloads"].items():
            hp = r["e2e"]["host_pack"]
            print("    ", n, "value %.2f e2e %.2f (share %.2f, host %.2f ms) label %.3f reduce %.3f step %.3f" % (r["value"] / 1e9, r["e2e"]["value"] / 1e9, hp["packed_share_of_batch"], hp["host_ms_per_step"], r["roofline"]["kernel_ms"], r["roofline"]["reduce_kernels_ms"], r["ms_per_step"]))
    except Exception as e:
        print(v, "ERR", e)
PY
