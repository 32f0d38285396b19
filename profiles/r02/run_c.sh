#!/bin/bash
mkdir -p gpurun_out/r02; python profiles/r02/dbg_gen.py 2>&1 | tail
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --workload tiny --also none > gpurun_out/r02/bench_2gpu_tiny.json 2> gpurun_out/r02/bench_2gpu_tiny.err; echo rc=$?
grep -v "^\*\|OMP" gpurun_out/r02/bench_2gpu_tiny.err | tail
