#!/bin/bash
# 2 GPUs: torchrun shape of the driver's scaling run; merged result checked against a single-GPU pass inside bench.py
mkdir -p gpurun_out/r02
nproc; nvidia-smi topo -m | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02/bench_2gpu_a.json 2> gpurun_out/r02/bench_2gpu_a.err; echo bench rc=$?
tail -40 gpurun_out/r02/bench_2gpu_a.err
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_multidevice.py -x -q -rP 2>&1 | grep -E "merge|passed|failed|rror" | head -40
