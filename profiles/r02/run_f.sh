#!/bin/bash
mkdir -p gpurun_out/r02
GS_DEBUG_MERGE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-fastq > gpurun_out/r02/bench_2gpu_d.json 2> gpurun_out/r02/bench_2gpu_d.err; echo bench rc=$?
grep -E "gs merge|PARITY|rank 0: " gpurun_out/r02/bench_2gpu_d.err | tail -40
