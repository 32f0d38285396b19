#!/bin/bash
# full GPU suite with the round-2 kernels (label diet, flat filter, L2 window, GSF1 files), then the whole default bench line
mkdir -p gpurun_out/r02
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 1200 python bench.py --steps 50 --warmup 5 > gpurun_out/r02/bench_o.json 2> gpurun_out/r02/bench_o.err; echo bench rc=$?
python - <<'PY'
import json
j = json.load(open("gpurun_out/r02/bench_o.json"))
print("viral value %.2f e2e %.2f label %.3f ms frac %.3f wall %.0f s" % (j["value"] / 1e9, j["e2e"]["value"] / 1e9, j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["bench_wall_s"]), j["cpu_baseline"].get("parity_kmers_per_taxon_equal"), "fastq %.2f" % (j["e2e_fastq"]["value"] / 1e9))
for n, r in j["workloads"].items():
    print(n, "value %.2f e2e %.2f kernel %.3f ms" % (r["value"] / 1e9, r["e2e"]["value"] / 1e9, r["roofline"]["kernel_ms"]), {k: v for k, v in (r.get("cpu_baseline") or {}).items() if k.startswith("parity") or k == "value"})
PY
