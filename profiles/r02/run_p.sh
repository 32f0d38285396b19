#!/bin/bash
# 2 GPUs: multi-rank / multi-device tests with the log kept, then the driver's torchrun shape of the bench (viral + bacterial + longread, merged result checked)
mkdir -p gpurun_out/r02
nvidia-smi topo -m | head -6
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_multidevice.py tests/test_gpu_match.py -q -rP 2>&1 | grep -E "merge [0-9]|passed|failed|rror|Error|skipped" | head -40 | tee gpurun_out/r02/pytest_gpu_2xB200.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r02/bench_2gpu_p.json 2> gpurun_out/r02/bench_2gpu_p.err; echo bench rc=$?
grep -E "PARITY|FAILED|Error" gpurun_out/r02/bench_2gpu_p.err | head
python - <<'PY'
import json
j = json.load(open("gpurun_out/r02/bench_2gpu_p.json"))
print("viral value %.2f e2e %.2f merge %.2f ms parity %s wall %.0f" % (j["value"] / 1e9, j["e2e"]["value"] / 1e9, j["end_of_job_reduce_ms"], j["merge"]["merge_parity"], j["bench_wall_s"]))
for n, r in j["workloads"].items():
    print(n, "value %.2f e2e %.2f merge %.2f ms parity %s" % (r["value"] / 1e9, r["e2e"]["value"] / 1e9, r["end_of_job_reduce_ms"], r["merge"]["merge_parity"]))
PY
