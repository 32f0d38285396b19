#!/bin/bash
# block-gzip input on the B200 box: the inflate kernel alone, then the match goal from a BGZF file (device inflate / host
# threads / zlib's sequential gzread) at two chunk sizes.  Results under gpurun_out/.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_inflate.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python profiles/microbench/inflate_bw.py 64 256 1024 > gpurun_out/inflate_bw.txt 2> gpurun_out/inflate_bw.err
cat gpurun_out/inflate_bw.txt
for mb in ${GS_CHUNKS:-64 256}; do
    GS_CHUNK_MB=$mb GS_BGZF=1 GS_SKIP_HOST=1 timeout 300 python profiles/microbench/goal_fastq.py ${GS_GOAL_READS:-16000000} > gpurun_out/goal_bgzf_chunk$mb.json 2> gpurun_out/goal_bgzf.err
    echo "chunk $mb MB rc=$?"
    python - "$mb" <<'PY'
import json, sys
d = json.load(open("gpurun_out/goal_bgzf_chunk%s.json" % sys.argv[1]))
for k, v in d.items():
    if isinstance(v, dict):
        print(" ", k, round(v["reads_per_s"] / 1e6, 2), "M reads/s", {kk: vv for kk, vv in v.items() if kk not in ("reads_per_s", "kmers_per_s")})
    else:
        print(" ", k, v)
PY
done
tail -3 gpurun_out/goal_bgzf.err
