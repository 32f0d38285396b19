#!/bin/bash
# usage: run_variants.sh "<variants>" "<extra bench flags>"
for v in $1; do
  for pf in "" "--no-prefilter"; do
    GS_LIB_VARIANT=/root/repo/genestrip_b200/_lib/$v.so python bench.py --no-cpu-baseline --steps 10 $pf $2 > gpurun_out/v.json 2> gpurun_out/v.err
    python - "$v" "$pf" <<PY
import json,sys
try:
    j=json.load(open("gpurun_out/v.json")); print(sys.argv[1], sys.argv[2] or "prefilter", "value %.2f G  e2e %.2f G  kernel_ms %.3f"%(j["value"]/1e9, j["e2e"]["value"]/1e9, j["roofline"]["kernel_ms"]))
except Exception as e: print(sys.argv[1], sys.argv[2], "ERR", e)
PY
  done
done
