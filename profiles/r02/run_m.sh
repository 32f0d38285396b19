#!/bin/bash
# flat filter kernel with probe order by L2 residency: parity, then the filter workload over resident size / rounds
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_goals.py -x -q -k "filter" 2>&1 | tail -3
for cfg in "0 0" "48 16" "32 12" "64 16" "48 27" "24 16" "48 8"; do
  set -- $cfg
  GS_FILTER_RESIDENT_MB=$1 GS_FILTER_RESIDENT_ROUNDS=$2 timeout 600 python bench.py --workload filter --also none --steps 20 --warmup 5 --cpu-seconds 3 > gpurun_out/r02/bench_m_$1_$2.json 2> gpurun_out/r02/bench_m_$1_$2.err; echo "MB=$1 rounds=$2 rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02/bench_m_*.json")):
    try:
        j = json.load(open(f))
        print(f, "value %.2f e2e %.2f ms %.3f" % (j["value"] / 1e9, j["e2e"]["value"] / 1e9, j["ms_per_step"]), j.get("native_options"), (j.get("cpu_baseline") or {}).get("parity_accept_bits_equal"), j.get("accepted_read_fraction"))
    except Exception as e:
        print(f, "ERR", e)
PY
