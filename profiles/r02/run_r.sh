#!/bin/bash
# N GPUs (argument), the driver's torchrun shape; viral + bacterial + longread with the merged result checked
N=${1:-4}
mkdir -p gpurun_out/r02
nproc; free -g | head -2 | tail -1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02/bench_${N}gpu_r.json 2> gpurun_out/r02/bench_${N}gpu_r.err; echo bench rc=$?
grep -E "PARITY|FAILED|Error|error" gpurun_out/r02/bench_${N}gpu_r.err | head
python - $N <<'PY'
import json, sys
N = sys.argv[1]
j = json.load(open("gpurun_out/r02/bench_%sgpu_r.json" % N))
print("viral value %.2f e2e %.2f fastq %.2f merge %.2f ms parity %s wall %.0f" % (j["value"] / 1e9, j["e2e"]["value"] / 1e9, (j.get("e2e_fastq") or {}).get("value", 0) / 1e9, j["end_of_job_reduce_ms"], j["merge"]["merge_parity"], j["bench_wall_s"]))
for n, r in j["workloads"].items():
    if "value" in r: print(n, "value %.2f e2e %.2f merge %.2f ms parity %s" % (r["value"] / 1e9, r["e2e"]["value"] / 1e9, r["end_of_job_reduce_ms"], r["merge"]["merge_parity"]))
    else: print(n, r)
PY
