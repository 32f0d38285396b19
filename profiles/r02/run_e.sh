#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_dbupdate.py -x -q 2>&1 | tail -3
GS_DEBUG_MERGE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload bacterial --also none --no-fastq > gpurun_out/r02/bench_2gpu_c.json 2> gpurun_out/r02/bench_2gpu_c.err; echo bench rc=$?
grep -E "gs merge|PARITY" gpurun_out/r02/bench_2gpu_c.err | tail -30
