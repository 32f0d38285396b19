#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_multidevice.py tests/test_gpu_dbupdate.py -x -q -rP 2>&1 | grep -E "merge [0-9]|passed|failed|rror|Error" | head -40
GS_DEBUG_MERGE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-fastq > gpurun_out/r02/bench_2gpu_e.json 2> gpurun_out/r02/bench_2gpu_e.err; echo bench rc=$?
grep -E "gs merge rank 0|PARITY|FAILED" gpurun_out/r02/bench_2gpu_e.err | tail -40
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02/bench_2gpu_f.json 2> gpurun_out/r02/bench_2gpu_f.err; echo bench rc=$?
