"""Throughput of gs_inflate_blocks (block-gzip members inflated on the device) on FASTQ text of the bench's shape.
usage: python profiles/microbench/inflate_bw.py [text_MB ...]   -> one JSON line per size"""
import json, os, sys, time, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from genestrip_b200 import capi
from concurrent.futures import ProcessPoolExecutor

sizes = [int(a) for a in sys.argv[1:]] or [64, 256, 1024]
rng = np.random.default_rng(3)
L = 150
n = max(sizes) * (1 << 20) // (2 * L + 16) + 1
seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(n, L))
qual = rng.choice(np.frombuffer(b"FFFFFFFF:,#", dtype=np.uint8), size=(n, L))   # Illumina-like binned qualities
rec = np.empty((n, 11 + L + 3 + L + 1), dtype=np.uint8)
rec[:, 0] = ord("@"); rec[:, 1] = ord("r")
idx = np.arange(n)
for d in range(9):
    rec[:, 2 + d] = idx // 10 ** (8 - d) % 10 + 48
rec[:, 10] = 10; rec[:, 11:11 + L] = seq; rec[:, 11 + L] = 10; rec[:, 12 + L] = ord("+"); rec[:, 13 + L] = 10
rec[:, 14 + L:14 + 2 * L] = qual; rec[:, -1] = 10
text_binned = rec.tobytes()
rec[:, 14 + L:14 + 2 * L] = ord("I")                                             # constant qualities: long runs (overlapping matches)
text_runs = rec.tobytes()
ctx = capi.Context([0])
for mb, (kind, text_all) in [(m, kt) for kt in (("binned random qualities", text_binned), ("constant qualities", text_runs)) for m in sizes]:
    text = text_all[:mb << 20]
    piece = 0xff00 * 64
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as pool:
        parts = list(pool.map(util.bgzf_bytes, [text[i:i + piece] for i in range(0, len(text), piece)], [0xff00] * ((len(text) + piece - 1) // piece),
                              [6] * ((len(text) + piece - 1) // piece), [False] * ((len(text) + piece - 1) // piece)))
    comp = b"".join(parts)
    blocks, total = capi.bgzf_blocks(comp)
    assert total == len(text)
    cbuf = capi.PinnedBuffer(len(comp)); cbuf.array[:len(comp)] = np.frombuffer(comp, dtype=np.uint8)
    obuf = capi.PinnedBuffer(total)
    best = None
    for rep in range(4):
        b = blocks.copy()
        t0 = time.perf_counter()
        rc = capi.lib().gs_inflate_blocks(ctx.h, cbuf.array.ctypes.data, len(comp), b.ctypes.data, len(b), obuf.array.ctypes.data, total)
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    assert obuf.array[:total].tobytes() == text
    print(json.dumps({"data": kind, "text_MB": mb, "blocks": len(blocks), "compressed_MB": round(len(comp) / 2 ** 20, 1), "ratio": round(total / len(comp), 2),
                      "seconds_incl_h2d_d2h": best, "text_GB_per_s": total / best / 1e9}))
