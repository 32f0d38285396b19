#!/bin/bash
# round 2, first GPU pass: parity of the packed link + the new bench line with its sub-workloads (1 GPU)
mkdir -p gpurun_out/r02
nproc; free -g | head -2; lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_match.py tests/test_gpu_scale.py -x -q 2>&1 | tail -5
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_a.json 2> gpurun_out/r02/bench_a.err; echo bench rc=$?
tail -30 gpurun_out/r02/bench_a.err
