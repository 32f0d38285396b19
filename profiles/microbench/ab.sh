#!/bin/bash
# usage: ab.sh "<bench flags common>"   -- prints value / e2e with and without the minimizer prefilter
for pf in "" "--no-prefilter"; do
  python bench.py --no-cpu-baseline --steps 10 $pf $1 > gpurun_out/v.json 2> gpurun_out/v.err
  python - "$pf" <<PY
import json,sys
try:
    j=json.load(open("gpurun_out/v.json")); print(sys.argv[1] or "prefilter", "value %.2f G  e2e %.2f G  kernel_ms %.3f h=%.3f"%(j["value"]/1e9, j["e2e"]["value"]/1e9, j["roofline"]["kernel_ms"], j["roofline"]["hit_fraction"]))
except Exception as e: print(sys.argv[1], "ERR", e); print(open("gpurun_out/v.err").read()[-2000:])
PY
done
