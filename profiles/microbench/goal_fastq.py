"""Goal-level throughput of `match` from a FASTQ FILE (parse + match + CSV), GPU feeder vs the sequential host parser.
usage: python profiles/microbench/goal_fastq.py [n_reads] [dir]   (viral-scale synthetic db as in bench.py)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from genestrip_b200 import capi, host

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
out_dir = sys.argv[2] if len(sys.argv) > 2 else "/dev/shm"
wl = dict(bench.WORKLOADS["viral"])
dev = torch.device("cuda", 0)
keys, vals_raw, parent, codes = bench.make_database(torch, dev, bench.DATABASES[wl["db"]], seed=43)
V = len(parent)
ctx = capi.Context([0])
db = capi.Database.from_pointers(ctx, bench.K, keys.data_ptr(), vals_raw.data_ptr(), keys.numel(), V, parent, build_bloom=True)
n_db = keys.numel()
del keys, vals_raw
L = wl["read_len"]
path = os.path.join(out_dir, "gs_goal_bench.fastq")
HDR = 11
rec_len = HDR + L + 3 + L + 1
with open(path, "wb") as f:
    done = 0
    while done < n_reads:
        R = min(2_000_000, n_reads - done)
        bases, _ = bench.make_reads(torch, dev, wl, codes, R, seed=777 + done)
        t = torch.empty((R, rec_len), dtype=torch.uint8, device=dev)
        t[:, 0] = ord("@"); t[:, 1] = ord("r")
        idx = torch.arange(done, done + R, device=dev)
        for d in range(9):
            t[:, 2 + d] = (idx // (10 ** (8 - d)) % 10 + 48).to(torch.uint8)
        t[:, HDR - 1] = 10
        t[:, HDR:HDR + L] = bases[:R * L].view(R, L)
        t[:, HDR + L] = 10; t[:, HDR + L + 1] = ord("+"); t[:, HDR + L + 2] = 10
        t[:, HDR + L + 3:HDR + 2 * L + 3] = ord("I")
        t[:, rec_len - 1] = 10
        f.write(t.view(-1).cpu().numpy().tobytes())
        done += R
        del t, bases
size = os.path.getsize(path)
# metadata for the CSV: synthetic names
taxids = [str(v + 1) for v in range(V)]
level = np.zeros(V, dtype=np.int64)
for v in range(1, V):
    level[v] = level[parent[v]] + 1
meta = host.DbMeta(bench.K, n_db, taxids, taxids, [-1] * V, parent, list(range(V)), level, [1] * V, [0] * V)
res = {}
chunk_mb = int(os.environ.get("GS_CHUNK_MB", "64"))
for name, kw in (("gpu_feeder", dict(gpu_parse=True, text_chunk_bytes=chunk_mb << 20)), ("host_parser", dict(gpu_parse=False, batch_reads=1 << 20))):
    if name == "host_parser" and os.environ.get("GS_SKIP_HOST"):
        continue
    best = None
    for rep in range(2):
        t0 = time.perf_counter()
        r = host.match_goal(db, meta, [path], **kw)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert r.total_reads == n_reads
    res[name] = {"seconds": best, "reads_per_s": n_reads / best, "kmers_per_s": r.total_kmers / best, "file_GB_per_s": size / best / 1e9,
                 "text_chunks": r.text_chunks, "csv_rows": r.csv.count(b"\n")}
    res[name + "_csv_md5"] = __import__("hashlib").md5(r.csv).hexdigest()
if os.environ.get("GS_BGZF"):
    # the same text as a block-gzip file: blocks inflated by the feeder's threads vs zlib's sequential gzread (GS_NO_BGZF)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from concurrent.futures import ProcessPoolExecutor
    gz_path = path + ".gz"
    piece = 0xff00 * 256
    with open(path, "rb") as f, open(gz_path, "wb") as g, ProcessPoolExecutor(max_workers=os.cpu_count()) as pool:
        parts = []
        while True:
            buf = f.read(piece)
            if not buf:
                break
            parts.append(pool.submit(util.bgzf_bytes, buf, 0xff00, 1, False))
        for fu in parts:
            g.write(fu.result())
        g.write(util.bgzf_bytes(b""))
    res["bgzf_file_bytes"] = os.path.getsize(gz_path)
    for name, env in (("gpu_feeder_bgzf_device_inflate", None), ("gpu_feeder_bgzf_host_threads", "GS_GPU_INFLATE=0"), ("gpu_feeder_gzread", "GS_NO_BGZF=1")):
        if env:
            os.environ[env.split("=")[0]] = env.split("=")[1]
        dt = None
        tm0 = host.bgzf_timers()
        for rep in range(1 if env and "NO_BGZF" in env else 2):
            t0 = time.perf_counter()
            r = host.match_goal(db, meta, [gz_path], gpu_parse=True, text_chunk_bytes=chunk_mb << 20)
            d = time.perf_counter() - t0
            dt = d if dt is None else min(dt, d)
        if env:
            os.environ.pop(env.split("=")[0], None)
        assert r.total_reads == n_reads
        res[name] = {"seconds": dt, "reads_per_s": n_reads / dt, "kmers_per_s": r.total_kmers / dt, "text_GB_per_s": size / dt / 1e9,
                     "inflate_call_s_all_reps": host.bgzf_timers()[0] - tm0[0], "inflate_calls_all_reps": host.bgzf_timers()[1] - tm0[1],
                     "bgzf_reader_s_all_reps": host.bgzf_timers()[2] - tm0[2],
                     "text_chunks": r.text_chunks, "same_csv_as_plain_file": __import__("hashlib").md5(r.csv).hexdigest() == res["gpu_feeder_csv_md5"]}
    res["inflate_threads"] = int(os.environ.get("GS_INFLATE_THREADS", "0")) or min(32, os.cpu_count())
    res["device_inflated_blocks"] = host.device_inflated_blocks()
    os.remove(gz_path)
res["same_csv"] = res["gpu_feeder_csv_md5"] == res.get("host_parser_csv_md5")
res["chunk_mb"] = chunk_mb
res["file_bytes"] = size
res["n_reads"] = n_reads
print(json.dumps(res))
os.remove(path)
