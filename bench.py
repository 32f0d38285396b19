#!/usr/bin/env python
"""bench.py -- `match` goal throughput (k-mers/s, reads/s) of the B200-native Genestrip hot path.

Contract (driver): python bench.py --gpus N --steps K --warmup W [--impl reference]; for N > 1 launched by torchrun,
one rank per GPU.  A step = one pass of the hot path (encode -> Bloom -> store lookup -> per-taxon counting ->
per-read classification) over one batch of synthetic reads.  Rank 0 prints ONE JSON line.

The line's own metric is quoted on BASELINE.json configs[1] (viral-scale database, 150 bp reads):
  value      k-mers/s, inputs resident in HBM when the timed region starts (device-resident C-ABI entry point), end-of-run
             merge across the ranks included (gs_match_finish_comm: NCCL all-reduce + peer-mapped OR/popcount kernel)
  e2e        the same metric through gs_match_submit / gs_match_collect with pinned HOST buffers (host packing, H2D, kernels,
             D2H inside the timed region)
  roofline   HBM-bound: algorithmic bytes per k-mer (DESIGN.md "Roofline") x k-mers per launch / launch duration
  cpu_baseline  the CPU oracle (restated reference algorithm, all host threads) on a bounded sample, rank 0, N = 1
  workloads  sub-records of the same shape for the other BASELINE.json configurations, measured in the same run:
             N = 1: bacterial (configs[2]), longread (configs[4]), filter (configs[3]); N > 1: bacterial, longread.
             Every sub-record carries its own parity flags (oracle spot check at N = 1; at N > 1 the merged multi-GPU result
             against a single-GPU pass over all ranks' batches).
"""
import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (default)
    "viral": dict(kind="match", db="viral", read_len=150, reads_per_step=4_000_000, frac_db=0.5, sub_rate=0.01, indel_rate=0.0,
                  desc="viral-scale synthetic db (~10k leaf taxa, 1e8 31-mers), match on 150 bp Illumina-like reads (BASELINE.json configs[1])"),
    "tiny": dict(kind="match", db="tiny", read_len=150, reads_per_step=200_000, frac_db=0.7, sub_rate=0.01, indel_rate=0.0,
                 desc="tiny synthetic db (9 leaf taxa, 2e6 31-mers), smoke-size"),
    # configs[2]: bacterial scale (database replicated per GPU), unique k-mer counting on
    "bacterial": dict(kind="match", db="bacterial", read_len=150, reads_per_step=4_000_000, frac_db=0.1, sub_rate=0.01, indel_rate=0.0,
                      desc="bacterial-scale synthetic db (50 625 leaf taxa, 2e9 31-mers), match with unique k-mer counting on 150 bp reads (BASELINE.json configs[2])"),
    # configs[4]: long reads, high hit rate: 1 % substitutions + 0.2 % single-base indels
    "longread": dict(kind="match", db="viral", read_len=10_000, reads_per_step=60_000, frac_db=0.9, sub_rate=0.01, indel_rate=0.002,
                     desc="long-read workload: 10 kb ONT-like reads, 90 % from the viral-scale db, 1 % substitutions + 0.2 % single-base indels (BASELINE.json configs[4])"),
    # configs[3]: the filter goal, XOR Bloom index (fpp 1e-8, 27 hashes) over the k-mers of the leaf taxa, ~1 % of the reads hit
    "filter": dict(kind="filter", db="viral", read_len=150, reads_per_step=4_000_000, frac_db=0.01, sub_rate=0.01, indel_rate=0.0,
                   desc="filter goal: XOR Bloom index (fpp 1e-8) of the viral-scale db, 150 bp reads with ~1 % hit rate (BASELINE.json configs[3])"),
}
DATABASES = {
    "viral": dict(levels=4, fanout=10, n_kmers=100_000_000),
    "tiny": dict(levels=2, fanout=3, n_kmers=2_000_000),
    "bacterial": dict(levels=4, fanout=15, n_kmers=2_000_000_000, simple_db=True),
}
P_PREFILTER_PASS = 0.09   # fraction of absent k-mers whose minimizer is in the store (measured, DESIGN.md "minimizer prefilter")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------------------------
# synthetic project on the GPU (torch is plumbing here: random numbers, sort, device memory)
# --------------------------------------------------------------------------------------------------------------
def make_database(torch, dev, wl, seed):
    """Genomes for every leaf of a fan-out tree; 1 % of every genome is copied from its next sibling so that LCA != leaf
    occurs.  Returns sorted distinct canonical 31-mers, raw Java-short value indices, the parent array, the genomes."""
    from genestrip_b200 import synth
    parent, level_start = synth.parent_array_fanout(wl["levels"], wl["fanout"])
    V = len(parent)
    leaf0 = level_start[-1]
    n_leaves = V - leaf0
    glen = int(math.ceil(wl["n_kmers"] / n_leaves * 1.004)) + K - 1
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    codes = torch.randint(0, 4, (n_leaves, glen), generator=g, device=dev, dtype=torch.int8)  # C0 G1 A2 T3
    if not wl.get("simple_db"):
        nsh = max(K, glen // 100)
        sib = (torch.arange(n_leaves, device=dev) + 1) % wl["fanout"] + (torch.arange(n_leaves, device=dev) // wl["fanout"]) * wl["fanout"]
        codes[:, glen // 2: glen // 2 + nsh] = codes[sib, glen // 3: glen // 3 + nsh]
    # canonical k-mers of every window (CGAT.java:145-265 semantics: fwd big-endian 2-bit, rc = reversed complement, max)
    nwin = glen - K + 1
    canon = torch.empty((n_leaves, nwin), dtype=torch.int64, device=dev)
    rows = max(1, (1 << 28) // nwin)                                    # <= 2 GiB temporaries per chunk
    for r0 in range(0, n_leaves, rows):
        cc = codes[r0:r0 + rows]
        fwd = torch.zeros((cc.shape[0], nwin), dtype=torch.int64, device=dev)
        rc = torch.zeros((cc.shape[0], nwin), dtype=torch.int64, device=dev)
        for i in range(K):
            c = cc[:, i:i + nwin].to(torch.int64)
            fwd = (fwd << 2) | c
            rc = rc | ((c ^ 1) << (2 * i))
        canon[r0:r0 + rows] = torch.maximum(fwd, rc)
        del fwd, rc
    canon = canon.reshape(-1)
    if wl.get("simple_db"):
        # bacterial scale: purely random genomes (duplicates are a handful), first holder wins; kept memory-lean
        keys, order = torch.sort(canon)
        del canon
        leaf32 = (order // nwin).to(torch.int32)
        del order
        keep = torch.ones(keys.numel(), dtype=torch.bool, device=dev)
        keep[1:] = keys[1:] != keys[:-1]
        keys = keys[keep]
        vals_raw = (leaf32[keep].to(torch.int32) + (level_start[-1] - 32768)).to(torch.int16)
        return keys, vals_raw, parent, codes
    leaf = (torch.arange(n_leaves, device=dev, dtype=torch.int64).repeat_interleave(nwin))
    keys, order = torch.sort(canon)
    leaf = leaf[order]
    del canon, order
    uniq, inv, cnt = torch.unique_consecutive(keys, return_inverse=True, return_counts=True)
    # value of a k-mer = LCA of all leaves holding it: longest common prefix of the leaf's base-`fanout` digits
    n_u = uniq.numel()
    lo = torch.full((n_u,), n_leaves, dtype=torch.int64, device=dev).scatter_reduce(0, inv, leaf, reduce="amin")
    hi = torch.zeros((n_u,), dtype=torch.int64, device=dev).scatter_reduce(0, inv, leaf, reduce="amax")
    vidx = torch.zeros((n_u,), dtype=torch.int64, device=dev)          # root
    found = torch.zeros((n_u,), dtype=torch.bool, device=dev)
    for lvl in range(wl["levels"], 0, -1):                               # deepest level where all holders agree
        div = wl["fanout"] ** (wl["levels"] - lvl)
        same = (lo // div == hi // div) & ~found
        vidx = torch.where(same, level_start[lvl] + lo // div, vidx)
        found |= same
    vals_raw = (vidx - 32768).to(torch.int16)
    return uniq, vals_raw, parent, codes


def make_reads(torch, dev, wl, codes, n_reads, seed):
    """n_reads x read_len ASCII bases on the device (+64 bytes of slack), offsets uint64[n+1].  frac_db of the reads are
    sampled from the genomes (random strand) with substitutions and, if indel_rate > 0, single-base insertions / deletions
    (half each); the others are iid random."""
    READ_LEN = wl["read_len"]
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n_leaves, glen = codes.shape
    indel = float(wl.get("indel_rate", 0.0))
    slack = 64 if indel > 0 else 0
    gi = torch.randint(0, n_leaves, (n_reads,), generator=g, device=dev)
    st = torch.randint(0, glen - READ_LEN + 1 - slack, (n_reads,), generator=g, device=dev)
    strand = torch.rand((n_reads,), generator=g, device=dev) < 0.5
    from_db = torch.rand((n_reads,), generator=g, device=dev) < wl["frac_db"]
    ar = torch.arange(READ_LEN, device=dev)
    rnd = torch.randint(0, 4, (n_reads, READ_LEN), generator=g, device=dev, dtype=torch.int8)
    if indel > 0:
        # walk the genome forward: a deletion skips a genome base, an insertion emits a random base without consuming one
        ev = torch.rand((n_reads, READ_LEN), generator=g, device=dev)
        ins = ev < indel / 2
        dele = (ev >= indel / 2) & (ev < indel)
        adv = (1 + dele.to(torch.int32) - ins.to(torch.int32))
        pos = torch.cumsum(adv, dim=1, dtype=torch.int32) - adv
        del ev, dele, adv
        pos = torch.clamp(st[:, None] + pos.to(torch.int64), max=glen - 1)
        c = codes[gi[:, None], pos]
        del pos
        c = torch.where(ins, rnd, c)
        del ins
        c = torch.where(strand[:, None], torch.flip(c, dims=(1,)) ^ 1, c)
    else:
        col = torch.where(strand[:, None], st[:, None] + (READ_LEN - 1 - ar)[None, :], st[:, None] + ar[None, :])
        c = codes[gi[:, None], col]
        c = torch.where(strand[:, None], c ^ 1, c)
    c = torch.where(from_db[:, None], c, rnd)
    sub = torch.rand((n_reads, READ_LEN), generator=g, device=dev) < wl["sub_rate"]
    c = torch.where(sub, rnd, c)
    lut = torch.tensor(list(b"CGAT"), dtype=torch.uint8, device=dev)
    bases = torch.zeros(n_reads * READ_LEN + 64, dtype=torch.uint8, device=dev)
    bases[: n_reads * READ_LEN] = lut[c.to(torch.int64)].reshape(-1)
    offsets = (torch.arange(n_reads + 1, device=dev, dtype=torch.int64) * READ_LEN)
    return bases, offsets


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (NVML; B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            log("clock sampling unavailable:", e)

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {getattr(nv, n): n for n in dir(nv) if n.startswith("nvmlClocksThrottleReason") or n.startswith("nvmlClocksEventReason")}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if isinstance(bit, int) and bit and (r & bit) and bit & (bit - 1) == 0:
                        self.reasons.add(nm.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", ""))
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        rs = sorted(x for x in self.reasons if x not in ("GpuIdle", "None", "All"))
        norm = {"SwPowerCap": "sw_power_cap", "HwSlowdown": "hw_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown",
                "SwThermalSlowdown": "sw_thermal_slowdown", "HwPowerBrakeSlowdown": "hw_power_brake_slowdown",
                "ApplicationsClocksSetting": "applications_clocks_setting", "SyncBoost": "sync_boost",
                "UserDefinedClocks": "applications_clocks_setting", "DisplayClockSetting": "display_clock_setting"}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(set(norm.get(r, r) for r in rs)), "samples": len(self.samples)}


def algorithmic_bytes_per_kmer(n_db, h, READ_LEN, use_bloom=True, count_unique=True, f=0.01):
    """SURVEY.md §8(d) / DESIGN.md "Roofline": ASCII base stream + two Bloom words + binary-search keys and the short value
    for the k-mers that reach the search + the unique-bitset RMW for hits."""
    s = h + (1.0 - h) * (f if use_bloom else 1.0)
    return (READ_LEN / (READ_LEN - K + 1)) + (16.0 if use_bloom else 0.0) + s * (8.0 * math.ceil(math.log2(max(n_db, 2))) + 2.0) + (8.0 * h if count_unique else 0.0)


def layout_bytes_per_kmer(h, READ_LEN):
    """What the device layout itself has to move per k-mer (DESIGN.md "Roofline"): the base, the 4-byte label written by the
    label kernel and read back by the reduce kernel, and one 32-byte probe-table sector for every k-mer that passes the
    L2-resident minimizer prefilter (all hits + ~9 % of the misses)."""
    return READ_LEN / (READ_LEN - K + 1) + 8.0 + 32.0 * (h + (1.0 - h) * P_PREFILTER_PASS)


def as_tensor(torch, ptr, nbytes, dev):
    """Wrap a raw device pointer (owned by the C library) as a uint8 torch tensor (no copy)."""
    class _Raw:
        pass
    r = _Raw()
    r.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(r, device=dev)


def java_random_longs(seed, n):
    """java.util.Random(seed).nextLong() sequence (hashFactors of AbstractKMerBloomFilter, C/bloom/AbstractKMerBloomFilter.java:104-110)."""
    mask = (1 << 48) - 1
    st = (seed ^ 0x5DEECE66D) & mask
    out = []
    for _ in range(n):
        parts = []
        for _ in range(2):
            st = (st * 0x5DEECE66D + 0xB) & mask
            v = st >> 16
            parts.append(v - (1 << 32) if v >= (1 << 31) else v)
        x = ((parts[0] << 32) + parts[1]) & ((1 << 64) - 1)
        out.append(x - (1 << 64) if x >= (1 << 63) else x)
    return out


def build_xor_index(torch, dev, keys, fpp):
    """The `filter` goal's index as BloomIndexGoal builds it (XORKMerBloomFilter, fpp 1e-8): bits, hashes, factors, words.
    torch is plumbing here (the index is built offline by the reference's db goals, outside the timed path)."""
    n = keys.numel()
    bits = max(1, int(-n * math.log(fpp) / (math.log(2.0) ** 2)))
    hashes = max(1, int(math.floor(bits / n * math.log(2.0) + 0.5)))
    factors = java_random_longs(42, hashes)
    n_words = (bits + 63) // 64
    flags = torch.zeros(n_words * 64, dtype=torch.bool, device=dev)
    for f in factors:
        idx = torch.fmod(keys ^ f, bits).abs()       # Math.abs((factor ^ key) % bits), Java's truncating remainder
        flags[idx] = True
    words = torch.empty(n_words, dtype=torch.int64, device=dev)
    weights = (torch.ones(64, dtype=torch.int64, device=dev) << torch.arange(64, device=dev))
    step = 1 << 24
    for w0 in range(0, n_words, step):
        blk = flags[w0 * 64:(w0 + step) * 64].view(-1, 64).to(torch.int64)
        words[w0:w0 + blk.shape[0]] = (blk * weights).sum(dim=1)
    return bits, hashes, np.array(factors, dtype=np.int64), words


def jvm_probe():
    """the CPU arm is the real reference when this box can run it (JDK >= 11 and the reference's jar built
    offline), else the C++ restatement.  Returns (usable, description)."""
    java = shutil.which("java")
    if not java:
        return False, "java: not found on PATH -> restated C++ port"
    try:
        out = subprocess.run([java, "-version"], capture_output=True, text=True, timeout=20)
        ver = (out.stderr or out.stdout).splitlines()[0]
    except Exception as e:
        return False, "java -version failed (%s) -> restated C++ port" % e
    jar = os.environ.get("GENESTRIP_JAR", "")
    if not jar or not os.path.exists(jar):
        return False, "%s present but no reference jar (set GENESTRIP_JAR; the image has no Maven repository to build it offline) -> restated C++ port" % ver
    return True, "%s, jar %s" % (ver, jar)


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name):
    """Per-launch DRAM traffic of the dominant kernel from the committed ncu --set full capture of this bench's own launch
    (profiles/r02/kernel_traffic.json, written by profiles/r02/extract_traffic.py); None when there is no capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02", "kernel_traffic.json"))).get(name)
    except Exception:
        return None


class Env:
    """Per-process plumbing: device, ranks, torch.distributed (barriers) and the library's own communicator (the merge)."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.comm = None
        self.ctx = None
        self.capi = None
        self.dbs = {}
        self.affinity0 = os.sched_getaffinity(0)
        self.threads = os.cpu_count() or 1
        self.t_start = time.perf_counter()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def database(self, name):
        """(keys, vals_raw, parent, codes) of a synthetic database, generated once per process (same seed on every rank)."""
        if name not in self.dbs:
            t0 = time.perf_counter()
            self.dbs[name] = make_database(self.torch, self.dev, DATABASES[name], seed=43)
            self.torch.cuda.synchronize()
            log("rank %d: database '%s': %d k-mers, %d tree nodes (%.1f s)" % (self.rank, name, self.dbs[name][0].numel(), len(self.dbs[name][2]), time.perf_counter() - t0))
        return self.dbs[name]

    def drop_database(self, name):
        self.dbs.pop(name, None)
        self.torch.cuda.empty_cache()


def workload_config(name, wl, n_db, V, world):
    READ_LEN, R = wl["read_len"], wl["reads_per_step"]
    return {"workload": "%s: %s" % (name, wl["desc"]), "k": K, "db_kmers": int(n_db), "tree_nodes": int(V), "read_len": READ_LEN,
            "reads_per_step": R, "kmers_per_step": R * (READ_LEN - K + 1), "frac_reads_from_db": wl["frac_db"], "substitution_rate": wl["sub_rate"],
            "indel_rate": wl.get("indel_rate", 0.0), "count_unique_kmers": True, "classify_reads": True, "use_bloom_filter": True,
            "l2_policy": "inputs larger than L2: %.0f MB of bases per step, %.0f MB database, alternating batches" % (R * READ_LEN / 1e6, n_db * 10 / 1e6),
            "parallelism": "reads sharded over %d GPU(s), database replicated" % world}


# --------------------------------------------------------------------------------------------------------------
# CPU legs (rank 0): the oracle as the reference's stand-in, and as the checker of the bench's own results
# --------------------------------------------------------------------------------------------------------------
def cpu_match(gs_oracle, odb, bases_h, offsets_h, threads, target_s, steps=1):
    """Times the CPU oracle (restated reference algorithm, reference threading model) on a bounded sample."""
    cfg = gs_oracle.match_cfg(k=K)
    n_all = len(offsets_h) - 1
    probe = min(n_all, max(200, int(2_000_000 // max(1, int(offsets_h[1] - offsets_h[0])))))
    t0 = time.perf_counter()
    odb.match_reads_mt(cfg, bases_h, offsets_h[: probe + 1], threads)
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(min(n_all, max(probe, probe * target_s / dt)))
    times, kmers, per = [], 0, None
    for _ in range(steps):
        t0 = time.perf_counter()
        kmers, per = odb.match_reads_mt(cfg, bases_h, offsets_h[: n + 1], threads)
        times.append(time.perf_counter() - t0)
    return n, kmers, per, times


def host_ram_ok(need_gb):
    try:
        import psutil
        return psutil.virtual_memory().available / 1e9 >= need_gb, psutil.virtual_memory().available / 1e9
    except Exception:
        return True, -1.0


# --------------------------------------------------------------------------------------------------------------
# match workloads
# --------------------------------------------------------------------------------------------------------------
def run_match(E, name, wl, want_fastq, want_cpu):
    torch, dev, capi, args, rank, world = E.torch, E.dev, E.capi, E.args, E.rank, E.world
    READ_LEN, R = wl["read_len"], wl["reads_per_step"]
    keys, vals_raw, parent, codes = E.database(wl["db"])
    V, n_db = len(parent), keys.numel()
    kmers_per_step = R * (READ_LEN - K + 1)
    n_batches = 2  # alternate between two resident batches (each far larger than the 126 MB L2)
    batches = [make_reads(torch, dev, wl, codes, R, seed=4343 + 1000 * rank + b) for b in range(n_batches)]
    torch.cuda.synchronize()
    config = workload_config(name, wl, n_db, V, world)
    db = capi.Database.from_pointers(E.ctx, K, keys.data_ptr(), vals_raw.data_ptr(), n_db, V, parent, build_bloom=True)
    log("rank %d: %s: database on device: %.2f GB, %d reads/step" % (rank, name, db.device_bytes / 1e9, R))
    cfg = capi.default_match_cfg(layout=capi.GS_LAYOUT_CLASSIC if args.layout == "classic" else capi.GS_LAYOUT_TABLE)
    cfg.prefilter = 0 if args.no_prefilter else 1
    cores_per_rank = max(1, len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    # two cores stay free for the submitting thread's own work and the driver's threads when the rank has eight or more: every
    # pack() is a barrier over the pool, and a pool as wide as the machine waits for whichever thread was descheduled
    # (measured: 14 threads 64.9-65.6 G k-mers/s e2e, 16 threads 63.5-63.7 G on the 16-core box)
    cfg.host_pack_threads = args.pack_threads if args.pack_threads >= 0 else (cores_per_rank - 2 if cores_per_rank >= 8 else cores_per_rank)
    cfg.host_pack_percent = args.pack_percent
    native_options = {"layout": args.layout, "minimizer_prefilter": bool(cfg.prefilter) and args.layout == "table", "host_pack_threads": int(cfg.host_pack_threads)}
    sess = capi.MatchSession(db, cfg)
    l2w = sess.l2_window()
    native_options["l2_persisting_window"] = ({"what": "minimizer prefilter", "window_bytes": l2w[0], "persisting_bytes": l2w[1], "hit_ratio": l2w[2]} if l2w[0] else None)
    if E.comm:
        sess.prepare_merge(E.comm)   # set-up, not part of the job: peer mappings of the bitsets, merge kernel loaded
    stream = torch.cuda.ExternalStream(sess.stream, device=dev)
    d_out = torch.zeros(R * 16, dtype=torch.uint8, device=dev)
    steps_done = args.warmup + args.steps

    def device_step(s, i, bb=batches):
        bases, offsets = bb[i % n_batches]
        s.run_device(bases.data_ptr(), offsets.data_ptr(), R, R * READ_LEN, i * R, d_out.data_ptr())

    for i in range(args.warmup):
        device_step(sess, i)
    sess.sync()
    sess.set_timing(True)   # CUDA events on the compute stream around the label kernel and the reduce kernels of every step
    launches0 = sess.kernel_launches
    sampler = ClockSampler(E.local)
    sampler.start()
    E.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(args.steps):
            device_step(sess, args.warmup + i)
            if i == args.steps - 1:
                sess.join()   # the last batch's reduce kernels run on the session's second stream: the closing event waits for them
            ev[i + 1].record(stream)
    sess.sync()
    E.barrier()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.result()
    label_ms, reduce_ms, _nb = sess.kernel_times()
    sess.set_timing(False)

    # ---------------- end of job: the merge across the ranks, inside the library (the only exchange step of the path)
    E.barrier()
    counts, _ = sess.finish(E.comm)
    launches = sess.kernel_launches - launches0
    merge_ms, merge_bits_ms, merge_bytes, merge_path = sess.merge_stats() if E.comm else (0.0, 0.0, 0, 0)
    sess.close()  # releases the probe table's in-line seen bits for the next session
    total_hits = int(counts["kmers"].sum())
    unique_total = int(counts["unique_kmers"].sum())

    # ---------------- N > 1: the merged result against ONE GPU doing every rank's batches (rank 0; the others wait)
    merge_parity = None
    if E.comm:
        if rank == 0:
            chk = capi.MatchSession(db, cfg)
            for r in range(world):
                bb = batches if r == 0 else [make_reads(torch, dev, wl, codes, R, seed=4343 + 1000 * r + b) for b in range(n_batches)]
                torch.cuda.synchronize()   # the session's stream does not wait for torch's: the reads must exist before the kernels start
                for i in range(steps_done):
                    device_step(chk, i, bb)
                chk.sync()
                del bb
            c1, _ = chk.finish()
            chk.close()
            fields = ("kmers", "contigs", "contig_len_squared_sum", "reads_1kmer", "reads", "reads_kmers", "reads_bps", "unique_kmers", "max_contig_len", "max_contig_read_no")
            bad = [f for f in fields if not np.array_equal(c1[f], counts[f])]
            merge_parity = not bad
            for f in bad:
                log("  %s: merged sum %d, single-GPU sum %d, %d of %d entries differ" % (f, int(counts[f].astype(np.int64).sum()), int(c1[f].astype(np.int64).sum()), int((counts[f] != c1[f]).sum()), len(c1[f])))
            if bad:
                log("MERGE PARITY FAILURE (%s): fields %s differ between the %d-GPU merge and the single-GPU pass" % (name, bad, world))
            torch.cuda.empty_cache()
        E.barrier()

    # ---------------- end-to-end: pinned host buffers through submit/collect (host packing + H2D + kernels + D2H per step)
    nb = R * READ_LEN
    pinned = [capi.PinnedBuffer(nb + 64) for _ in range(n_batches)]
    pin_off = capi.PinnedBuffer((R + 1) * 8)
    for b in range(n_batches):
        pinned[b].array[:nb] = batches[b][0][:nb].cpu().numpy()
    offs = pin_off.view(np.uint64, R + 1)
    off_h = batches[0][1].cpu().numpy().astype(np.uint64)
    offs[:] = off_h
    sess2 = capi.MatchSession(db, cfg)
    e2e_seen = [0]  # per-read result records received on the host (zero-copy views of the pinned staging buffers)

    def e2e_run(n_steps, first):
        pend = []
        for i in range(n_steps):
            pend.append(sess2.submit(pinned[(first + i) % n_batches].array, offs, (first + i) * R))
            if len(pend) == capi.GS_MAX_INFLIGHT:
                e2e_seen[0] += len(sess2.collect_view(pend.pop(0))[0])
        while pend:
            e2e_seen[0] += len(sess2.collect_view(pend.pop(0))[0])

    # the link split of gs_match_submit settles over the first batches of a session (a real run has thousands): the e2e leg
    # warms up a little longer than the kernels need
    e2e_warmup = max(args.warmup, 16) if cfg.host_pack_threads != 0 and cfg.host_pack_percent < 0 else args.warmup
    e2e_run(e2e_warmup, 0)
    p0 = sess2.pack_stats()
    E.barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps, e2e_warmup)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    E.barrier()
    p1 = sess2.pack_stats()
    pack_frac = sess2.pack_fraction
    pack_threads, pack_s, h2d_base_bytes = p1[0], p1[1] - p0[1], (p1[3] - p0[3]) / max(1, args.steps)

    # ---------------- end-to-end from raw FASTQ text (GPU feeder: the device splits the records; no host parsing at all)
    fq_s, fq_bytes = 0.0, 0
    if READ_LEN <= 1000 and want_fastq:
        HDR = 11  # "@r%09d\n"
        rec_len = HDR + READ_LEN + 3 + READ_LEN + 1
        fq_bytes = R * rec_len
        pinned_fq = []
        for b in range(n_batches):
            t = torch.empty((R, rec_len), dtype=torch.uint8, device=dev)
            t[:, 0] = ord("@"); t[:, 1] = ord("r")
            idx = torch.arange(R, device=dev)
            for d in range(9):
                t[:, 2 + d] = (idx // (10 ** (8 - d)) % 10 + 48).to(torch.uint8)
            t[:, HDR - 1] = 10
            t[:, HDR:HDR + READ_LEN] = batches[b][0][:nb].view(R, READ_LEN)
            t[:, HDR + READ_LEN] = 10; t[:, HDR + READ_LEN + 1] = ord("+"); t[:, HDR + READ_LEN + 2] = 10
            t[:, HDR + READ_LEN + 3:HDR + 2 * READ_LEN + 3] = ord("I")
            t[:, rec_len - 1] = 10
            pb = capi.PinnedBuffer(fq_bytes + 64)
            pb.array[:fq_bytes] = t.view(-1).cpu().numpy()
            pinned_fq.append(pb)
            del t
        torch.cuda.empty_cache()
        fq_reads = [0]

        def fq_run(n_steps, first):
            pend = []
            for i in range(n_steps):
                tk, info = sess2.submit_fastq(pinned_fq[(first + i) % n_batches].array, (first + i) * R, n_bytes=fq_bytes)
                assert tk and info.n_reads == R, "FASTQ feeder refused the synthetic chunk (status %d)" % info.status
                pend.append(tk)
                if len(pend) == capi.GS_MAX_INFLIGHT:
                    fq_reads[0] += len(sess2.collect_fastq(pend.pop(0))[0])
            while pend:
                fq_reads[0] += len(sess2.collect_fastq(pend.pop(0))[0])

        fq_run(args.warmup, 0)
        E.barrier()
        t0 = time.perf_counter()
        fq_run(args.steps, args.warmup)
        torch.cuda.synchronize()
        fq_s = time.perf_counter() - t0
        E.barrier()
        for pb in pinned_fq:
            pb.free()
    sess2.close()

    # max over ranks of the timed regions
    total_ms_max, e2e_ms_max, fq_ms_max, merge_ms_max = E.max_over_ranks([total_ms + merge_ms, e2e_s * 1e3, fq_s * 1e3, merge_ms])
    h = total_hits / float(steps_done * kmers_per_step * world)
    value = world * args.steps * kmers_per_step / (total_ms_max / 1e3)
    e2e_value = world * args.steps * kmers_per_step / (e2e_ms_max / 1e3)

    rec = None
    if rank == 0:
        peak, peak_src = peak_hbm()
        bpk = algorithmic_bytes_per_kmer(n_db, h, READ_LEN)
        step_mean = float(np.mean(step_ms))
        # the dominant kernel: gs_label_kernel does everything B_kmer counts (bases, filter words, store search, unique bits);
        # the reduce kernels only re-read its 4-byte labels and write the 16-byte per-read records
        kernel_ms = label_ms if label_ms > 0 else step_mean
        achieved = bpk * kmers_per_step / (kernel_ms / 1e3) / 1e9
        lbpk = layout_bytes_per_kmer(h, READ_LEN)
        tj = ncu_traffic(name) if (args.layout == "table" and not args.no_prefilter) else None
        traffic = None
        if tj and tj.get("reads_per_launch") == R:
            traffic = tj["dram_bytes_per_launch"]
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "gs_label_kernel<%s>" % args.layout, "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / step_mean,
                "reduce_kernels_ms": reduce_ms, "step_ms": step_mean, "step_frac": bpk * kmers_per_step / (step_mean / 1e3) / 1e9 / peak,
                "timing": "CUDA events on the session's compute stream around the kernel, every timed step (gs_match_kernel_times)",
                "algorithmic_bytes_per_kmer": bpk, "hit_fraction": h, "peak_source": peak_src,
                "layout_bytes_per_kmer": lbpk,
                "layout_frac": lbpk * kmers_per_step / (kernel_ms / 1e3) / 1e9 / peak,
                "note": "frac prices SURVEY.md §8(d)'s algorithmic bytes (the reference's %d-step binary search + two Bloom words); the device layout answers a k-mer with at most one 32-byte sector behind an L2-resident prefilter (layout_bytes_per_kmer), so frac can exceed 1 -- layout_frac and dram_frac are the fractions of the HBM peak this layout really uses" % math.ceil(math.log2(max(n_db, 2))),
                "traffic_source": "profiles/r02/kernel_traffic.json: ncu --set full of this bench's own launch (dram__bytes_read.sum + dram__bytes_write.sum)" if traffic else None}
        if traffic:
            roof["dram_frac"] = traffic / (kernel_ms / 1e3) / 1e9 / peak
            roof["dram_bytes_per_kmer"] = traffic / kmers_per_step
        if tj and "dram_lines_per_kmer" in tj:
            lps = kmers_per_step / (kernel_ms / 1e3) * tj["dram_lines_per_kmer"]
            roof["request_roofline"] = {"measured_cap_lines_per_s": 39.4e9, "dram_lines_per_kmer": tj["dram_lines_per_kmer"], "lines_per_s": lps, "frac": lps / 39.4e9,
                                        "l2_hit_rate_pct": tj.get("l2_hit_rate_pct"),
                                        "source": "profiles/microbench/randgroup.txt: divergent loads are capped at ~39.4 G distinct 128-byte lines/s"}
        rec = {"metric": "match k-mers/s", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
               "data": "synthetic", "reads_per_s": value / (READ_LEN - K + 1), "config": config, "native_options": native_options, "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": int(h2d_base_bytes) + (R + 1) * 8, "d2h_bytes_per_step": R * 16 + 8 + V * 16,
                       "reads_per_s": e2e_value / (READ_LEN - K + 1), "warmup": e2e_warmup,
                       "host_pack": {"threads": pack_threads, "packed_share_of_batch": pack_frac, "isa": capi.lib().gs_pack_isa().decode(), "host_ms_per_step": pack_s * 1e3 / max(1, args.steps),
                                     "ascii_bytes_per_step": nb,
                                     "what": "gs_match_submit splits every batch: the tail crosses the link as ASCII while the host threads pack the head to 2-bit codes + validity bits (0.375 bytes per base), inside the timed region; the split follows the measured cost of the two routes"} if pack_threads else None},
               "gpu_launches": int(launches),
               "roofline": roof,
               "e2e_fastq": ({"value": world * args.steps * kmers_per_step / (fq_ms_max / 1e3), "unit": "k-mers/s",
                              "reads_per_s": world * args.steps * R / (fq_ms_max / 1e3), "h2d_bytes_per_step": fq_bytes,
                              "d2h_bytes_per_step": R * 32 + 16 + V * 20,
                              "what": "gs_match_submit_fastq / gs_match_collect_fastq: raw 4-line FASTQ text in pinned host memory, records split on the GPU"}
                             if fq_ms_max > 0 else None),
               "end_of_job_reduce_ms": merge_ms_max, "hits_total": total_hits, "unique_kmers_total": unique_total}
        if E.comm:
            rec["merge"] = {"where": "gs_match_finish_comm (library, NCCL loaded at run time)", "total_ms": merge_ms_max, "bitset_ms": merge_bits_ms,
                            "bitset_bytes_read_from_peers_per_rank": int(merge_bytes), "bitset_bytes_per_rank": int(merge_bytes) * world // max(1, world - 1),
                            "path": {1: "peer mappings (cudaIpc) read in place by the OR/popcount kernel", 2: "ncclSend/ncclRecv slice exchange"}.get(merge_path, "none"),
                            "merge_parity": merge_parity,
                            "merge_parity_what": "merged counters, unique k-mers and max-contigs of the %d ranks == one GPU processing every rank's %d batches" % (world, steps_done)}

    # ---------------- CPU baseline + bench-scale parity spot check (rank 0, N = 1 only)
    if rank == 0 and world == 1 and want_cpu:
        import gs_oracle
        os.sched_setaffinity(0, E.affinity0)
        need = n_db * 22 / 1e9 * 2.2 + 8
        ok, avail = host_ram_ok(need)
        if not ok:
            rec["cpu_baseline"] = {"skipped": "host RAM: %.0f GB available, %.0f GB needed for the oracle's copy of the %d-key store" % (avail, need, n_db)}
            log("CPU BASELINE SKIPPED for %s: %s" % (name, rec["cpu_baseline"]["skipped"]))
        else:
            t0 = time.perf_counter()
            keys_h, vals_h = keys.cpu().numpy(), vals_raw.cpu().numpy()
            odb = gs_oracle.OracleDb.from_arrays(K, keys_h, vals_h, V, parent, build_bloom=True)
            del keys_h, vals_h
            build_s = time.perf_counter() - t0
            b0_h = pinned[0].array[:nb]
            cpu_s = min(args.cpu_seconds, 8.0) if READ_LEN > 1000 else args.cpu_seconds
            n, kmers, per, times = cpu_match(gs_oracle, odb, b0_h, off_h, E.threads, cpu_s)
            odb.free()
            sess3 = capi.MatchSession(db, cfg)
            t = sess3.submit(b0_h, off_h[: n + 1].copy(), 0)
            sess3.collect(t)
            c3, _ = sess3.finish()
            sess3.close()
            parity = bool(np.array_equal(c3["kmers"], per))
            rec["cpu_baseline"] = {"value": kmers / times[0], "unit": "k-mers/s", "cores": E.threads, "kind": "port",
                                   "sample": "first %d reads of one step (%.1f s of CPU work; oracle store built in %.0f s); C++ restatement of FastqKMerMatcher.matchRead on the full %d-key store, %d matcher threads" % (n, times[0], build_s, n_db, E.threads),
                                   "parity_kmers_per_taxon_equal": parity}
            if not parity:
                log("PARITY FAILURE on the bench sample (%s)" % name)
    for p in pinned:
        p.free()
    pin_off.free()
    db.close()
    del batches, d_out
    torch.cuda.empty_cache()
    return rec


# --------------------------------------------------------------------------------------------------------------
# filter workload (no reduction at all: a pure map over the reads)
# --------------------------------------------------------------------------------------------------------------
def run_filter(E, name, wl, want_cpu):
    """BASELINE.json configs[3]: FastqBloomFilter.isAcceptRead over the XOR index; same JSON contract, metric = filter k-mers/s."""
    torch, dev, capi, args, rank, world = E.torch, E.dev, E.capi, E.args, E.rank, E.world
    READ_LEN, R = wl["read_len"], wl["reads_per_step"]
    keys, vals_raw, parent, codes = E.database(wl["db"])
    dbp = DATABASES[wl["db"]]
    V, n_db = len(parent), keys.numel()
    n_batches = 2
    batches = [make_reads(torch, dev, wl, codes, R, seed=4545 + 1000 * rank + b) for b in range(n_batches)]
    config = workload_config(name, wl, n_db, V, world)
    leaf0 = len(parent) - dbp["fanout"] ** dbp["levels"]
    leaf_keys = keys[(vals_raw.to(torch.int64) + 32768) >= leaf0]   # index = k-mers of the requested (leaf) taxa
    bits, hashes, factors, words = build_xor_index(torch, dev, leaf_keys, 1e-8)
    words_h = words.cpu().numpy()
    flt = capi.Filter(E.ctx, capi.GS_BLOOM_XOR, bits, hashes, factors, words_h)
    n_index = int(leaf_keys.numel())
    del words, leaf_keys
    torch.cuda.empty_cache()
    log("rank %d: XOR index: %d keys, %d bits, %d hashes" % (rank, n_index, bits, hashes))
    sess = capi.FilterSession(flt, K, 1, 0.2)
    stream = torch.cuda.ExternalStream(sess.stream, device=dev)
    d_acc = torch.zeros(R, dtype=torch.uint8, device=dev)

    def step(i):
        b, o = batches[i % n_batches]
        sess.run_device(b.data_ptr(), o.data_ptr(), R, R * READ_LEN, d_acc.data_ptr())

    for i in range(args.warmup):
        step(i)
    sess.sync()
    sampler = ClockSampler(E.local)
    sampler.start()
    launches0 = sess.kernel_launches
    E.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
            ev[i + 1].record(stream)
    sess.sync()
    E.barrier()
    launches = sess.kernel_launches - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.result()
    accepted = int(d_acc.sum().item())
    # end to end
    nb = R * READ_LEN
    pinned = [capi.PinnedBuffer(nb + 64) for _ in range(n_batches)]
    pin_off = capi.PinnedBuffer((R + 1) * 8)
    for b in range(n_batches):
        pinned[b].array[:nb] = batches[b][0][:nb].cpu().numpy()
    offs = pin_off.view(np.uint64, R + 1)
    off_h = batches[0][1].cpu().numpy().astype(np.uint64)
    offs[:] = off_h

    def e2e_run(n_steps, first):
        pend = []
        for i in range(n_steps):
            pend.append(sess.submit(pinned[(first + i) % n_batches].array, offs))
            if len(pend) == capi.GS_MAX_INFLIGHT:
                sess.collect(pend.pop(0))
        while pend:
            sess.collect(pend.pop(0))

    e2e_run(args.warmup, 0)
    E.barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps, args.warmup)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    total_ms_max, e2e_ms_max = E.max_over_ranks([total_ms, e2e_s * 1e3])
    kmers_per_step = R * (READ_LEN - K + 1)
    value = world * args.steps * kmers_per_step / (total_ms_max / 1e3)
    e2e_value = world * args.steps * kmers_per_step / (e2e_ms_max / 1e3)
    rec = None
    if rank == 0:
        peak, peak_src = peak_hbm()
        hfrac = accepted / float(R)
        probes = (1 - hfrac * 0.73) * 2 + hfrac * 0.73 * hashes   # E[probes] ~ (1-h)*2 + h*27
        bpk = READ_LEN / (READ_LEN - K + 1) + 8.0 * probes
        kernel_ms = total_ms_max / args.steps
        achieved = bpk * kmers_per_step / (kernel_ms / 1e3) / 1e9
        tj = ncu_traffic(name)
        traffic = tj["dram_bytes_per_launch"] if tj and tj.get("reads_per_launch") == R else None
        dram, reqroof = None, None
        if traffic:  # every 8-byte Bloom probe drags a whole sector / line out of DRAM
            dram = {"achieved_dram_GBs": traffic / (kernel_ms / 1e3) / 1e9, "frac_of_peak": traffic / (kernel_ms / 1e3) / 1e9 / peak,
                    "dram_bytes_per_kmer": traffic / kmers_per_step,
                    "note": "real DRAM bytes per second: the kernel is bound by the lines its probes drag in, not by the 8 algorithmic bytes per probe"}
        if tj and "dram_lines_per_kmer" in tj:
            lps = kmers_per_step / (kernel_ms / 1e3) * tj["dram_lines_per_kmer"]
            reqroof = {"measured_cap_lines_per_s": 39.4e9, "dram_lines_per_kmer": tj["dram_lines_per_kmer"], "lines_per_s": lps, "frac": lps / 39.4e9,
                       "l2_hit_rate_pct": tj.get("l2_hit_rate_pct"), "source": "profiles/microbench/randgroup.txt"}
        rec = {"metric": "filter k-mers/s", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
               "reads_per_s": value / (READ_LEN - K + 1), "config": config, "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": nb + (R + 1) * 8, "d2h_bytes_per_step": R},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                            "kernel": "gs_filter_flat_kernel (+ gs_mark_starts_kernel, gs_filter_accept_kernel)", "kernel_ms": kernel_ms, "algorithmic_bytes_per_kmer": bpk, "peak_source": peak_src,
                            "dram": dram, "request_roofline": reqroof},
               "accepted_read_fraction": hfrac, "index": {"kind": "xor", "keys": n_index, "bits": bits, "hashes": hashes}}
    if rank == 0 and world == 1 and want_cpu:
        import gs_oracle
        os.sched_setaffinity(0, E.affinity0)
        oflt = gs_oracle.Bloom.from_words(1, bits, hashes, factors, words_h)
        b0_h = pinned[0].array[:nb]
        probe = 20000
        t0 = time.perf_counter()
        oflt.accept_reads_mt(K, b0_h, off_h[: probe + 1], E.threads)
        dt = max(time.perf_counter() - t0, 1e-6)
        n = int(min(R, max(probe, probe * args.cpu_seconds / dt)))
        t0 = time.perf_counter()
        kmers, acc_o = oflt.accept_reads_mt(K, b0_h, off_h[: n + 1], E.threads)
        cpu_s = time.perf_counter() - t0
        oflt.free()
        t = sess.submit(b0_h, off_h[: n + 1].copy())
        acc_g = sess.collect(t)
        parity = bool(np.array_equal(acc_g[:n], acc_o))
        rec["cpu_baseline"] = {"value": kmers / cpu_s, "unit": "k-mers/s", "cores": E.threads, "kind": "port",
                               "sample": "first %d reads of one step (%.1f s of CPU work); C++ restatement of FastqBloomFilter.isAcceptRead over the same XOR index, %d threads" % (n, cpu_s, E.threads),
                               "parity_accept_bits_equal": parity, "accepted_in_sample": int(acc_o.sum())}
        if not parity:
            log("PARITY FAILURE on the filter bench sample")
    sess.close()
    flt.close()
    for p in pinned:
        p.free()
    pin_off.free()
    del batches, d_acc
    torch.cuda.empty_cache()
    return rec


# --------------------------------------------------------------------------------------------------------------
def reference_arm(args, emit):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores (rank 0 only)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = args.gpus
    name = args.workload
    wl = dict(WORKLOADS[name])
    if args.reads_per_step:
        wl["reads_per_step"] = args.reads_per_step
    if wl["kind"] != "match":
        raise SystemExit("--impl reference times the match path")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py generates its synthetic project on a CUDA device")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    usable, probe = jvm_probe()
    log("reference arm: %s" % probe)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gs_oracle
    gs_oracle.build()
    keys, vals_raw, parent, codes = make_database(torch, dev, DATABASES[wl["db"]], seed=43)
    V, n_db = len(parent), keys.numel()
    R, READ_LEN = wl["reads_per_step"], wl["read_len"]
    bases, offsets = make_reads(torch, dev, wl, codes, R, seed=4343)
    config = workload_config(name, wl, n_db, V, world)
    b0_h = bases[: R * READ_LEN].cpu().numpy()
    off_h = offsets.cpu().numpy().astype(np.uint64)
    threads = os.cpu_count() or 1
    odb = gs_oracle.OracleDb.from_arrays(K, keys.cpu().numpy(), vals_raw.cpu().numpy(), V, parent, build_bloom=True)
    del keys, vals_raw, codes, bases
    torch.cuda.empty_cache()
    # every step is a bounded sample of the workload, sized so that the whole --steps K --warmup W run ends within a few minutes
    per_step_s = min(args.cpu_seconds, max(1.0, 150.0 / max(1, args.steps + args.warmup)))
    n, kmers, per, times = cpu_match(gs_oracle, odb, b0_h, off_h, threads, per_step_s, steps=max(1, args.steps + args.warmup))
    odb.free()
    times = times[args.warmup:] if len(times) > args.warmup else times
    dt = float(np.mean(times))
    val = kmers / dt
    emit({"impl": "reference", "metric": "match k-mers/s", "value": val, "unit": "k-mers/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
          "reads_per_s": n / dt, "config": config, "jvm_probe": probe,
          "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": threads, "kind": "port",
                           "sample": "%d of the %d reads of one step per timed step; C++ restatement of FastqKMerMatcher.matchRead with the reference's threading model (%s)" % (n, R, probe)},
          "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return 0


def main():
    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library chatter)
    # is sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GS_BENCH_WORKLOAD", "viral"), choices=sorted(WORKLOADS))
    ap.add_argument("--also", default=os.environ.get("GS_BENCH_ALSO", "auto"),
                    help="comma-separated workloads measured as sub-records ('none'; 'auto' = bacterial,longread,filter at N = 1 and bacterial,longread at N > 1 when the main workload is viral)")
    ap.add_argument("--reads-per-step", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layout", default="table", choices=["table", "classic"],
                    help="device index: 128-byte probe table (default) or the reference's Bloom filter + sorted-array search")
    ap.add_argument("--no-prefilter", action="store_true", help="A/B: probe the table for every k-mer (no minimizer prefilter)")
    ap.add_argument("--no-fastq", action="store_true", help="skip the raw-FASTQ end-to-end leg")
    ap.add_argument("--pack-threads", type=int, default=-1, help="host threads that pack the bases in gs_match_submit (0 = ASCII on the link; default: CPUs / ranks)")
    ap.add_argument("--pack-percent", type=int, default=-1, help="share of a batch that is packed (-1: adaptive)")
    ap.add_argument("--budget-seconds", type=float, default=600.0, help="sub-workloads are skipped once the run is this old")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args, emit)
    args.warmup = max(args.warmup, 3)

    E = Env(args)
    torch = E.torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(E.local)
    E.dev = torch.device("cuda", E.local)
    # run on the cores next to this rank's GPU so that the pinned batches are allocated on its NUMA node (restored for the CPU leg)
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(E.local)
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
        cpus = {i for i in range(os.cpu_count()) if (words[i // 64] >> (i % 64)) & 1} & E.affinity0
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:  # no NVML / no topology information: keep the default placement
        log("rank %d: cpu affinity not set (%s)" % (E.rank, e))
    from genestrip_b200 import capi
    from genestrip_b200 import dist as gsdist
    E.capi = capi
    E.ctx = capi.Context([E.local])
    if E.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=E.dev)
        E.dist = dist
        # torch.distributed is plumbing (barriers, max over ranks, handing out the communicator id); the merge itself runs in
        # the library over its own NCCL communicator.  Communicator set-up is not part of the job.
        w = torch.zeros(E.world * 4, dtype=torch.int64, device=E.dev)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        E.comm = gsdist.open_comm(dist, capi, E.ctx)
        torch.cuda.synchronize()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if E.rank == 0 and E.world == 1 and not args.no_cpu_baseline:
        import gs_oracle
        gs_oracle.build()

    def run(name, primary):
        wl = dict(WORKLOADS[name])
        if args.reads_per_step and primary:
            wl["reads_per_step"] = args.reads_per_step
        want_cpu = not args.no_cpu_baseline
        if wl["kind"] == "filter":
            return run_filter(E, name, wl, want_cpu)
        return run_match(E, name, wl, want_fastq=(primary and not args.no_fastq), want_cpu=want_cpu)

    line = run(args.workload, True)
    also = args.also
    if also == "auto":
        also = ("bacterial,longread,filter" if E.world == 1 else "bacterial,longread") if args.workload == "viral" else "none"
    names = [n for n in also.split(",") if n and n != "none"]
    subs = {}
    # cheap ones first; the database of the main workload is reused where the sub-workload names the same one
    order = sorted(names, key=lambda n: (WORKLOADS[n]["db"] != WORKLOADS[args.workload]["db"], n))
    for n in order:
        if n not in WORKLOADS or n == args.workload:
            continue
        # every rank takes the same decision (the sub-workloads contain collectives)
        elapsed = E.max_over_ranks([time.perf_counter() - E.t_start])[0]
        if elapsed > args.budget_seconds:
            subs[n] = {"skipped": "run older than --budget-seconds (%.0f s) before this workload" % args.budget_seconds}
            continue
        if WORKLOADS[n]["db"] != WORKLOADS[args.workload]["db"]:
            E.drop_database(WORKLOADS[args.workload]["db"])
        try:
            t0 = time.perf_counter()
            subs[n] = run(n, False)
            if E.rank == 0 and subs[n] is not None:
                subs[n]["wall_s"] = time.perf_counter() - t0
        except Exception as e:  # a failed sub-workload must not cost the headline line -- but it must be loud
            import traceback
            log("SUB-WORKLOAD %s FAILED:\n%s" % (n, traceback.format_exc()))
            subs[n] = {"error": "%s: %s" % (type(e).__name__, e)}
            if E.dist:   # ranks may be out of step now: no further collectives
                break
    if E.rank == 0:
        if names:
            line["workloads"] = subs
        line["bench_wall_s"] = time.perf_counter() - E.t_start
        emit(line)
    if E.comm:
        E.comm.close()
    E.ctx.close()
    if E.dist:
        E.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
