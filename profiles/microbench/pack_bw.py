"""Host-side packer throughput (gs_pack_bases) by thread count, pageable vs pinned input, no GPU activity.
usage: python profiles/microbench/pack_bw.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from genestrip_b200 import capi
L = capi.lib()
n = 600_000_000
rng = np.random.default_rng(1)
src = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)]
codes = np.zeros((n + 31) // 32 + 2, dtype=np.uint64)
valid = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
ctx = capi.Context([0])
pin = capi.PinnedBuffer(n + 64)
pin.array[:n] = src
pc = capi.PinnedBuffer(codes.nbytes); pv = capi.PinnedBuffer(valid.nbytes)
print("isa", L.gs_pack_isa().decode(), "cpus", len(os.sched_getaffinity(0)))
for name, b, c, v in (("pageable", src, codes, valid), ("pinned", pin.array, pc.view(np.uint64, len(codes)), pv.view(np.uint32, len(valid)))):
    for t in (1, 2, 4, 8, 12, 16, 24, 32):
        if t > 2 * len(os.sched_getaffinity(0)):
            break
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter()
            rc = L.gs_pack_bases(b.ctypes.data, n, c.ctypes.data, v.ctypes.data, t)
            best = min(best, time.perf_counter() - t0)
        print("%s threads %2d: %.2f ms  %.1f GB/s of ASCII" % (name, t, best * 1e3, n / best / 1e9), flush=True)
