#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q -rP 2>&1 | grep -E "merge [0-9]|passed|failed|rror|Error" | head -60
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02/bench_2gpu_b.json 2> gpurun_out/r02/bench_2gpu_b.err; echo bench rc=$?
grep -v "^\*\|OMP" gpurun_out/r02/bench_2gpu_b.err | tail -30
