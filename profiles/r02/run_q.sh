#!/bin/bash
# final 1-GPU lines: the default bench (viral + sub-records) with the refreshed ncu traffic file, the tiny config, the reference arm
mkdir -p gpurun_out/r02
timeout 1200 python bench.py > gpurun_out/r02/bench_q.json 2> gpurun_out/r02/bench_q.err; echo bench rc=$?
timeout 600 python bench.py --workload tiny --also none > gpurun_out/r02/bench_q_tiny.json 2> gpurun_out/r02/bench_q_tiny.err; echo tiny rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench_q_ref.json 2> gpurun_out/r02/bench_q_ref.err; echo ref rc=$?
python - <<'PY'
import json
j = json.load(open("gpurun_out/r02/bench_q.json"))
print("viral value %.2f e2e %.2f label %.3f ms frac %.3f dram_frac %.3f wall %.0f s" % (j["value"] / 1e9, j["e2e"]["value"] / 1e9, j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["roofline"].get("dram_frac") or 0, j["bench_wall_s"]), j["cpu_baseline"].get("parity_kmers_per_taxon_equal"), j["steps"])
for n, r in j["workloads"].items():
    print(n, "value %.2f e2e %.2f kernel %.3f ms frac %.3f" % (r["value"] / 1e9, r["e2e"]["value"] / 1e9, r["roofline"]["kernel_ms"], r["roofline"]["frac"]), {k: v for k, v in (r.get("cpu_baseline") or {}).items() if k.startswith("parity") or k == "value"})
t = json.load(open("gpurun_out/r02/bench_q_tiny.json"))
print("tiny value %.2f e2e %.2f" % (t["value"] / 1e9, t["e2e"]["value"] / 1e9), t["cpu_baseline"]["value"] / 1e6, t["cpu_baseline"].get("parity_kmers_per_taxon_equal"), t["roofline"]["frac"])
r = json.load(open("gpurun_out/r02/bench_q_ref.json"))
print("reference arm %.1f M k-mers/s" % (r["value"] / 1e6), r["jvm_probe"], r["config"] == j["config"])
PY
