#!/usr/bin/env python
"""bench.py -- `match` goal throughput (k-mers/s, reads/s) of the B200-native Genestrip hot path.

Contract (driver): python bench.py --gpus N --steps K --warmup W [--impl reference]; for N > 1 launched by torchrun,
one rank per GPU.  A step = one pass of the hot path (encode -> Bloom -> store lookup -> per-taxon counting ->
per-read classification) over one batch of synthetic 150 bp reads against the viral-scale synthetic database
(BASELINE.json configs[1]).  Rank 0 prints ONE JSON line.

  value      k-mers/s, inputs resident in HBM when the timed region starts (device-resident C-ABI entry point)
  e2e        the same metric through gs_match_submit/gs_match_collect with pinned HOST buffers (H2D + D2H inside)
  roofline   HBM-bound: algorithmic bytes per k-mer (DESIGN.md "Roofline") x k-mers per launch / launch duration
  cpu_baseline  the CPU oracle (restated reference algorithm, all host threads) on a bounded sample, rank 0, N=1
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (default)
    "viral": dict(kind="match", levels=4, fanout=10, n_kmers=100_000_000, read_len=150, reads_per_step=4_000_000, frac_db=0.5, sub_rate=0.01,
                  desc="viral-scale synthetic db (~10k leaf taxa, 1e8 31-mers), match on 150 bp Illumina-like reads (BASELINE.json configs[1])"),
    "tiny": dict(kind="match", levels=2, fanout=3, n_kmers=2_000_000, read_len=150, reads_per_step=200_000, frac_db=0.7, sub_rate=0.01,
                 desc="tiny synthetic db (9 leaf taxa, 2e6 31-mers), smoke-size"),
    # configs[2]: bacterial scale (database replicated per GPU), unique k-mer counting on
    "bacterial": dict(kind="match", levels=4, fanout=15, n_kmers=2_000_000_000, read_len=150, reads_per_step=4_000_000, frac_db=0.1, sub_rate=0.01,
                      simple_db=True, desc="bacterial-scale synthetic db (50 625 leaf taxa, 2e9 31-mers), match with unique k-mer counting on 150 bp reads (BASELINE.json configs[2])"),
    # configs[4]: long reads, high hit rate (substitutions only; indels are not generated)
    "longread": dict(kind="match", levels=4, fanout=10, n_kmers=100_000_000, read_len=10_000, reads_per_step=60_000, frac_db=0.9, sub_rate=0.01,
                     desc="long-read workload: 10 kb ONT-like reads, 90 % from the viral-scale db, 1 % substitutions (BASELINE.json configs[4])"),
    # configs[3]: the filter goal, XOR Bloom index (fpp 1e-8, 27 hashes) over the k-mers of the leaf taxa, ~1 % of the reads hit
    "filter": dict(kind="filter", levels=4, fanout=10, n_kmers=100_000_000, read_len=150, reads_per_step=4_000_000, frac_db=0.01, sub_rate=0.01,
                   desc="filter goal: XOR Bloom index (fpp 1e-8) of the viral-scale db, 150 bp reads with ~1 % hit rate (BASELINE.json configs[3])"),
}


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
    """n_reads x read_len ASCII bases on the device (+64 bytes of slack), offsets uint64[n+1]."""
    READ_LEN = wl["read_len"]
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n_leaves, glen = codes.shape
    gi = torch.randint(0, n_leaves, (n_reads,), generator=g, device=dev)
    st = torch.randint(0, glen - READ_LEN + 1, (n_reads,), generator=g, device=dev)
    strand = torch.rand((n_reads,), generator=g, device=dev) < 0.5
    from_db = torch.rand((n_reads,), generator=g, device=dev) < wl["frac_db"]
    ar = torch.arange(READ_LEN, device=dev)
    col = torch.where(strand[:, None], st[:, None] + (READ_LEN - 1 - ar)[None, :], st[:, None] + ar[None, :])
    c = codes[gi[:, None], col]
    c = torch.where(strand[:, None], c ^ 1, c)
    rnd = torch.randint(0, 4, (n_reads, READ_LEN), generator=g, device=dev, dtype=torch.int8)
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
            time.sleep(0.05)

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


def cpu_reference(gs_oracle, keys_h, vals_h, V, parent, bases_h, offsets_h, threads, target_s, steps=1):
    """Times the CPU oracle (restated reference algorithm, reference threading model) on a bounded sample."""
    odb = gs_oracle.OracleDb.from_arrays(K, keys_h, vals_h, V, parent, build_bloom=True)
    cfg = gs_oracle.match_cfg(k=K)
    n_all = len(offsets_h) - 1
    probe = min(n_all, 20000)
    t0 = time.perf_counter()
    odb.match_reads_mt(cfg, bases_h, offsets_h[: probe + 1], threads)
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(min(n_all, max(probe, probe * target_s / dt)))
    times, kmers, per = [], 0, None
    for _ in range(steps):
        t0 = time.perf_counter()
        kmers, per = odb.match_reads_mt(cfg, bases_h, offsets_h[: n + 1], threads)
        times.append(time.perf_counter() - t0)
    odb.free()
    return n, kmers, per, times


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


def run_filter_workload(torch, dev, args, wl, config, capi, ctx, keys, vals_raw, parent, batches, b0_h, off_h, emit, rank, world, dist):
    """BASELINE.json configs[3]: FastqBloomFilter.isAcceptRead over the XOR index; same JSON contract, metric = filter k-mers/s."""
    READ_LEN, R = wl["read_len"], wl["reads_per_step"]
    leaf0 = len(parent) - wl["fanout"] ** wl["levels"]
    leaf_keys = keys[(vals_raw.to(torch.int64) + 32768) >= leaf0]   # index = k-mers of the requested (leaf) taxa
    bits, hashes, factors, words = build_xor_index(torch, dev, leaf_keys, 1e-8)
    flt = capi.Filter(ctx, capi.GS_BLOOM_XOR, bits, hashes, factors, words.cpu().numpy())
    del words, keys, vals_raw
    log("rank %d: XOR index: %d keys, %d bits, %d hashes" % (rank, leaf_keys.numel(), bits, hashes))
    sess = capi.FilterSession(flt, K, 1, 0.2)
    stream = torch.cuda.ExternalStream(sess.stream, device=dev)
    d_acc = torch.zeros(R, dtype=torch.uint8, device=dev)
    n_batches = len(batches)

    def step(i):
        b, o = batches[i % n_batches]
        sess.run_device(b.data_ptr(), o.data_ptr(), R, d_acc.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sess.sync()
    sampler = ClockSampler(dev.index)
    sampler.start()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
            ev[i + 1].record(stream)
    sess.sync()
    barrier()
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
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps, args.warmup)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    tt = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    kmers_per_step = R * (READ_LEN - K + 1)
    value = world * args.steps * kmers_per_step / (float(tt[0]) / 1e3)
    e2e_value = world * args.steps * kmers_per_step / (float(tt[1]) / 1e3)
    if rank == 0:
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = 6650.0
        hfrac = accepted / float(R)
        probes = (1 - hfrac * 0.73) * 2 + hfrac * 0.73 * hashes   # E[probes] ~ (1-h)*2 + h*27
        bpk = READ_LEN / (READ_LEN - K + 1) + 8.0 * probes
        kernel_ms = float(tt[0]) / args.steps
        achieved = bpk * kmers_per_step / (kernel_ms / 1e3) / 1e9
        traffic, dram, reqroof = None, None, None
        try:  # ncu capture of the filter kernel: every 8-byte Bloom probe drags a 128-byte line out of DRAM
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01", "filter_kernel_traffic.json")))
            traffic = tj["dram_bytes_per_kmer"] * kmers_per_step
            dram = {"achieved_dram_GBs": traffic / (kernel_ms / 1e3) / 1e9, "frac_of_peak": traffic / (kernel_ms / 1e3) / 1e9 / peak,
                    "note": "real DRAM bytes per second: the kernel is bound by the lines its probes drag in, not by the 8 algorithmic bytes per probe"}
            lps = kmers_per_step / (kernel_ms / 1e3) * tj["dram_lines_per_kmer"]
            reqroof = {"measured_cap_lines_per_s": 39.4e9, "dram_lines_per_kmer": tj["dram_lines_per_kmer"], "lines_per_s": lps, "frac": lps / 39.4e9,
                       "source": "profiles/microbench/randgroup.txt"}
        except Exception:
            pass
        emit({"metric": "filter k-mers/s", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
              "reads_per_s": value / (READ_LEN - K + 1), "config": config, "clocks": clocks,
              "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": nb + (R + 1) * 8, "d2h_bytes_per_step": R},
              "gpu_launches": args.steps,
              "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                           "kernel": "gs_filter_kernel", "kernel_ms": kernel_ms, "algorithmic_bytes_per_kmer": bpk,
                           "dram": dram, "request_roofline": reqroof},
              "accepted_read_fraction": hfrac, "index": {"kind": "xor", "bits": bits, "hashes": hashes}})
    sess.close()
    flt.close()
    for p in pinned:
        p.free()
    pin_off.free()
    ctx.close()
    if dist:
        dist.destroy_process_group()
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GS_BENCH_WORKLOAD", "viral"), choices=sorted(WORKLOADS))
    ap.add_argument("--reads-per-step", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layout", default="table", choices=["table", "classic"],
                    help="device index: 128-byte probe table (default) or the reference's Bloom filter + sorted-array search")
    ap.add_argument("--no-prefilter", action="store_true", help="A/B: probe the table for every k-mer (no minimizer prefilter)")
    ap.add_argument("--no-fastq", action="store_true", help="skip the raw-FASTQ end-to-end leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    wl = dict(WORKLOADS[args.workload])
    if args.reads_per_step:
        wl["reads_per_step"] = args.reads_per_step
    READ_LEN = wl["read_len"]
    if args.workload == "bacterial":
        args.no_cpu_baseline = True   # a 2e9-key oracle database does not fit the bounded CPU leg; parity at this size: test_gpu_scale

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # run on the cores next to this rank's GPU so that the pinned batches are allocated on its NUMA node (restored for the CPU leg)
    affinity0 = os.sched_getaffinity(0)
    if args.impl == "native":
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
            cpus = {i for i in range(os.cpu_count()) if (words[i // 64] >> (i % 64)) & 1} & affinity0
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception as e:  # no NVML / no topology information: keep the default placement
            log("rank %d: cpu affinity not set (%s)" % (rank, e))
    dist = None
    if world > 1 and args.impl == "native":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        # communicator set-up is not part of the job: run every collective of the end-of-job merge once
        w = torch.zeros(world * 4, dtype=torch.int64, device=dev)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        dist.all_to_all_single(torch.empty_like(w), w)
        torch.cuda.synchronize()

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    threads = os.cpu_count() or 1

    # ---------------- synthetic project (same DB on every rank, different reads per rank)
    t0 = time.perf_counter()
    keys, vals_raw, parent, codes = make_database(torch, dev, wl, seed=43)
    V = len(parent)
    n_db = keys.numel()
    R = wl["reads_per_step"]
    n_batches = 2  # alternate between two resident batches (each far larger than the 126 MB L2)
    batches = [make_reads(torch, dev, wl, codes, R, seed=4343 + 1000 * rank + b) for b in range(n_batches)]
    torch.cuda.synchronize()
    log("rank %d: synthetic project: %d db k-mers, %d tree nodes, %d reads/step (%.1f s)" % (rank, n_db, V, R, time.perf_counter() - t0))
    kmers_per_step = R * (READ_LEN - K + 1)
    config = {"workload": "%s: %s" % (args.workload, wl["desc"]), "k": K, "db_kmers": int(n_db), "tree_nodes": int(V), "read_len": READ_LEN,
              "reads_per_step": R, "kmers_per_step": kmers_per_step, "frac_reads_from_db": wl["frac_db"], "substitution_rate": wl["sub_rate"],
              "count_unique_kmers": True, "classify_reads": True, "use_bloom_filter": True,
              "l2_policy": "inputs larger than L2: %.0f MB of bases per step, %.0f MB database, alternating batches" % (R * READ_LEN / 1e6, n_db * 10 / 1e6),
              "parallelism": "reads sharded over %d GPU(s), database replicated" % world}

    need_host_db = args.impl == "reference" or (rank == 0 and world == 1 and not args.no_cpu_baseline and wl["kind"] == "match")
    keys_h = keys.cpu().numpy() if need_host_db else None
    vals_h = vals_raw.cpu().numpy() if need_host_db else None
    b0_h = batches[0][0][: R * READ_LEN].cpu().numpy()
    off_h = batches[0][1].cpu().numpy().astype(np.uint64)

    if args.impl == "reference":
        import gs_oracle
        n, kmers, per, times = cpu_reference(gs_oracle, keys_h, vals_h, V, parent, b0_h, off_h, threads, args.cpu_seconds, steps=max(1, args.steps + args.warmup))
        times = times[args.warmup:] if len(times) > args.warmup else times
        dt = float(np.mean(times))
        val = kmers / dt
        line = {"impl": "reference", "metric": "match k-mers/s", "value": val, "unit": "k-mers/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "reads_per_s": n / dt, "config": config,
                "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": threads, "kind": "port",
                                 "sample": "%d of the %d reads of one step per timed step; C++ restatement of FastqKMerMatcher.matchRead with the reference's threading model (no JDK on this image)" % (n, R)},
                "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ---------------- native arm
    from genestrip_b200 import capi
    ctx = capi.Context([local])
    if wl["kind"] == "filter":
        return run_filter_workload(torch, dev, args, wl, config, capi, ctx, keys, vals_raw, parent, batches, b0_h, off_h, emit, rank, world, dist)
    torch.cuda.synchronize()
    db = capi.Database.from_pointers(ctx, K, keys.data_ptr(), vals_raw.data_ptr(), n_db, V, parent, build_bloom=True)
    del keys, vals_raw
    torch.cuda.empty_cache()
    log("rank %d: database on device: %.2f GB" % (rank, db.device_bytes / 1e9))
    cfg = capi.default_match_cfg(layout=capi.GS_LAYOUT_CLASSIC if args.layout == "classic" else capi.GS_LAYOUT_TABLE)
    cfg.prefilter = 0 if args.no_prefilter else 1
    config["layout"] = args.layout
    config["minimizer_prefilter"] = bool(cfg.prefilter) and args.layout == "table"
    sess = capi.MatchSession(db, cfg)
    stream = torch.cuda.ExternalStream(sess.stream, device=dev)
    d_out = torch.zeros(R * 16, dtype=torch.uint8, device=dev)

    def device_step(i):
        bases, offsets = batches[i % n_batches]
        sess.run_device(bases.data_ptr(), offsets.data_ptr(), R, R * READ_LEN, i * R, d_out.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        device_step(i)
    sess.sync()
    sess.set_timing(True)   # CUDA events on the compute stream around the label kernel and the reduce kernels of every step
    launches0 = sess.kernel_launches
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(args.steps):
            device_step(args.warmup + i)
            ev[i + 1].record(stream)
    sess.sync()
    barrier()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.result()
    launches = sess.kernel_launches - launches0
    label_ms, reduce_ms, _nb = sess.kernel_times()
    sess.set_timing(False)

    # ---------------- end of job: merge the per-rank state (only exchange step of the path)
    red_ms = 0.0
    counts, _ = None, None
    if dist:
        c_ptr, m_ptr, b_ptr, b_words = sess.device_state()
        t_c = as_tensor(torch, c_ptr, 7 * V * 8, dev).view(torch.int64)
        t_m = as_tensor(torch, m_ptr, V * 8, dev).view(torch.int64)
        t_b = as_tensor(torch, b_ptr, b_words * 8, dev).view(torch.int64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        from genestrip_b200.dist import merge_match_state

        def popcount_slice(merged, lo_w, hi_w):
            # the popcount kernel addresses bitset words absolutely: put the OR-merged slice back at its place
            uniq = torch.zeros(V, dtype=torch.int64, device=dev)
            t_b.zero_()
            if hi_w > lo_w:
                t_b[lo_w:hi_w] = merged
            torch.cuda.synchronize()
            sess.unique_popcount(b_ptr, lo_w, hi_w, uniq.data_ptr())
            sess.sync()
            return uniq

        uniq = merge_match_state(dist, t_c, t_m, t_b, V, popcount_slice)
        e1.record()
        torch.cuda.synchronize()
        red_ms = e0.elapsed_time(e1)
        total_hits = int(t_c[:V].sum().item())
        unique_total = int(uniq.sum().item())
    else:
        counts, _ = sess.finish()
        total_hits = int(counts["kmers"].sum())
        unique_total = int(counts["unique_kmers"].sum())
    sess.close()  # releases the probe table's in-line seen bits for the next session


    # ---------------- end-to-end: pinned host buffers through submit/collect (H2D + kernels + D2H per step)
    nb = R * READ_LEN
    pinned = [capi.PinnedBuffer(nb + 64) for _ in range(n_batches)]
    pin_off = capi.PinnedBuffer((R + 1) * 8)
    for b in range(n_batches):
        pinned[b].array[:nb] = batches[b][0][:nb].cpu().numpy()
    offs = pin_off.view(np.uint64, R + 1)
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

    e2e_run(args.warmup, 0)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps, args.warmup)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---------------- end-to-end from raw FASTQ text (GPU feeder: the device splits the records; no host parsing at all)
    fq_s, fq_bytes = 0.0, 0
    if READ_LEN <= 1000 and not args.no_fastq:
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
        barrier()
        t0 = time.perf_counter()
        fq_run(args.steps, args.warmup)
        torch.cuda.synchronize()
        fq_s = time.perf_counter() - t0
        barrier()
        for pb in pinned_fq:
            pb.free()

    # max over ranks of the timed region
    tt = torch.tensor([total_ms + red_ms, e2e_s * 1e3, fq_s * 1e3], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max, fq_ms_max = float(tt[0]), float(tt[1]), float(tt[2])
    steps_done = args.warmup + args.steps
    h = total_hits / float(steps_done * kmers_per_step * world) if dist else total_hits / float(steps_done * kmers_per_step)
    value = world * args.steps * kmers_per_step / (total_ms_max / 1e3)
    e2e_value = world * args.steps * kmers_per_step / (e2e_ms_max / 1e3)

    line = None
    if rank == 0:
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bpk = algorithmic_bytes_per_kmer(n_db, h, READ_LEN)
        traffic, tj = None, None
        try:  # DRAM bytes of the label kernel from the committed ncu --set full capture, scaled to this launch's k-mers
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01", "label_kernel_traffic.json")))
            if args.workload == "viral" and args.layout == "table" and bool(cfg.prefilter) == bool(tj.get("minimizer_prefilter", True)):
                traffic = tj["dram_bytes_per_kmer"] * kmers_per_step
        except Exception:
            tj = None
        step_ms = float(np.mean(step_ms))
        # the dominant kernel: gs_label_kernel does everything B_kmer counts (bases, filter words, store search, unique bits);
        # the reduce kernels only re-read its 4-byte labels and write the 16-byte per-read records
        kernel_ms = label_ms if label_ms > 0 else step_ms
        achieved = bpk * kmers_per_step / (kernel_ms / 1e3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "gs_label_kernel<%s>" % args.layout, "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / step_ms,
                "reduce_kernels_ms": reduce_ms, "step_ms": step_ms, "step_frac": bpk * kmers_per_step / (step_ms / 1e3) / 1e9 / peak,
                "timing": "CUDA events on the session's compute stream around the kernel, every timed step (gs_match_kernel_times)",
                "algorithmic_bytes_per_kmer": bpk, "hit_fraction": h, "peak_source": peak_src,
                "traffic_source": "profiles/r01/label_kernel_traffic.json (ncu --set full: dram__bytes_read+write per k-mer x k-mers per launch)" if traffic else None}
        if traffic and tj and "dram_lines_per_kmer" in tj:
            lps = kmers_per_step / (kernel_ms / 1e3) * tj["dram_lines_per_kmer"]
            roof["request_roofline"] = {"measured_cap_lines_per_s": 39.4e9, "dram_lines_per_kmer": tj["dram_lines_per_kmer"], "lines_per_s": lps, "frac": lps / 39.4e9,
                                        "source": "profiles/microbench/randgroup.txt: divergent loads are capped at ~39.4 G distinct 128-byte lines/s"}
        line = {"metric": "match k-mers/s", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
                "data": "synthetic", "reads_per_s": value / (READ_LEN - K + 1), "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": nb + (R + 1) * 8, "d2h_bytes_per_step": R * 16 + 4 + V * 16,
                        "reads_per_s": e2e_value / (READ_LEN - K + 1)},
                "gpu_launches": int(launches),
                "roofline": roof,
                "e2e_fastq": ({"value": world * args.steps * kmers_per_step / (fq_ms_max / 1e3), "unit": "k-mers/s",
                               "reads_per_s": world * args.steps * R / (fq_ms_max / 1e3), "h2d_bytes_per_step": fq_bytes,
                               "d2h_bytes_per_step": R * 32 + 16 + V * 20,
                               "what": "gs_match_submit_fastq / gs_match_collect_fastq: raw 4-line FASTQ text in pinned host memory, records split on the GPU"}
                              if fq_ms_max > 0 else None),
                "end_of_job_reduce_ms": red_ms, "hits_total": total_hits, "unique_kmers_total": unique_total}

    # ---------------- CPU baseline + bench-scale parity spot check (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import gs_oracle
        os.sched_setaffinity(0, affinity0)
        if READ_LEN > 1000:
            args.cpu_seconds = min(args.cpu_seconds, 8.0)
        n, kmers, per, times = cpu_reference(gs_oracle, keys_h, vals_h, V, parent, b0_h, off_h, threads, args.cpu_seconds)
        sess2.close()
        sess3 = capi.MatchSession(db, cfg)
        t = sess3.submit(b0_h, off_h[: n + 1].copy(), 0)
        sess3.collect(t)
        c3, _ = sess3.finish()
        sess3.close()
        parity = bool(np.array_equal(c3["kmers"], per))
        line["cpu_baseline"] = {"value": kmers / times[0], "unit": "k-mers/s", "cores": threads, "kind": "port",
                                "sample": "first %d reads of one step (%.1f s of CPU work); C++ restatement of FastqKMerMatcher.matchRead, %d matcher threads" % (n, times[0], threads),
                                "parity_kmers_per_taxon_equal": parity}
        if not parity:
            log("PARITY FAILURE on the bench sample")
    if rank == 0:
        emit(line)
    sess2.close()
    for p in pinned:
        p.free()
    pin_off.free()
    db.close()
    ctx.close()
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
