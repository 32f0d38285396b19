#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_match.py -x -q 2>&1 | tail -3
timeout 1200 python bench.py --steps 20 --warmup 5 --also longread > gpurun_out/r02/bench_i1.json 2> gpurun_out/r02/bench_i1.err; echo bench rc=$?
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --also none > gpurun_out/r02/bench_i2.json 2> gpurun_out/r02/bench_i2.err; echo bench rc=$?
for pct in 70 80 90; do timeout 600 python bench.py --steps 20 --warmup 5 --also none --no-fastq --no-cpu-baseline --pack-percent $pct > gpurun_out/r02/bench_i1_p$pct.json 2>> gpurun_out/r02/bench_i1.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02/bench_i*.json")):
    try:
        j=json.load(open(f)); hp=j["e2e"]["host_pack"] or {}
        print(f, "value %.1f e2e %.1f"%(j["value"]/1e9, j["e2e"]["value"]/1e9), hp.get("packed_share_of_batch"), hp.get("host_ms_per_step"), j["e2e"]["h2d_bytes_per_step"])
        for n,r in (j.get("workloads") or {}).items():
            hp=r["e2e"]["host_pack"] or {}
            print("   ",n, "value %.1f e2e %.1f"%(r["value"]/1e9, r["e2e"]["value"]/1e9), hp.get("packed_share_of_batch"), hp.get("host_ms_per_step"))
    except Exception as e: print(f, e)
PY
