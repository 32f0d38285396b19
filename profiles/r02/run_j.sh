#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 --also bacterial --no-fastq > gpurun_out/r02/bench_j1.json 2> gpurun_out/r02/bench_j1.err; echo bench rc=$?
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02/bench_j1.json"))
print("viral value %.2f e2e %.2f label %.3f ms"%(j["value"]/1e9, j["e2e"]["value"]/1e9, j["roofline"]["kernel_ms"]), j["cpu_baseline"]["parity_kmers_per_taxon_equal"])
for n,r in j["workloads"].items(): print(n, "value %.2f e2e %.2f label %.3f ms"%(r["value"]/1e9, r["e2e"]["value"]/1e9, r["roofline"]["kernel_ms"]), r.get("cpu_baseline"))
PY
bash profiles/r02/run_ncu.sh viral bacterial longread filter
