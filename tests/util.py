"""Shared helpers of the parity tests: build the same database in the oracle and on the GPU, run both."""
import numpy as np

from genestrip_b200 import synth


def build_pair(oracle, native, ctx, k, nodes, names, genomes, requested=(), device_bloom=False, **okw):
    """Oracle DB via the restated `db` goal, then the same arrays uploaded through the C ABI."""
    odb = oracle.OracleDb.build(k, nodes, names, genomes, requested=requested, **okw)
    gdb = upload(oracle, native, ctx, odb, device_bloom=device_bloom)
    return odb, gdb


def upload(oracle, native, ctx, odb, device_bloom=False, with_bloom=True):
    keys, vals = odb.export()
    parent, depth, position, has_node = odb.tree()
    bloom = None
    f = odb.store_filter()
    if with_bloom and not device_bloom and f is not None and f.kind == 0:
        _, seed, buckets, _, words = f.params()
        bloom = (seed, buckets, words)
    return native.Database(ctx, odb.k, keys, vals, odb.n_values, parent_by_vidx=parent, has_node=has_node,
                           bloom=bloom, build_bloom=("return" if (device_bloom and with_bloom) else False))


def gpu_match(native, gdb, bases, offsets, batch=None, first_read_no=0, **cfg):
    """Run the C-ABI match path over host buffers in batches; returns (results, events, counts, top, runs per batch)."""
    mcfg = native.default_match_cfg(**cfg)
    sess = native.MatchSession(gdb, mcfg)
    n = len(offsets) - 1
    batch = batch or max(n, 1)
    results, events, runs = [], [], []
    pending = []
    try:
        for b0 in range(0, max(n, 1), batch):
            b1 = min(n, b0 + batch)
            off = np.ascontiguousarray(offsets[b0:b1 + 1])
            t = sess.submit(bases, off, first_read_no + b0)
            pending.append(t)
            if len(pending) == native.GS_MAX_INFLIGHT:
                r, e, ro, ru = sess.collect(pending.pop(0))
                results.append(r); events.append(e); runs.append((ro, ru))
        while pending:
            r, e, ro, ru = sess.collect(pending.pop(0))
            results.append(r); events.append(e); runs.append((ro, ru))
        counts, top = sess.finish()
        launches = sess.kernel_launches
    finally:
        sess.close()
    res = np.concatenate(results) if results else np.zeros(0, dtype=native.READ_RESULT_DTYPE)
    ev = np.concatenate(events) if events else np.zeros(0, dtype=native.EVENT_DTYPE)
    return res, ev, counts, top, runs, launches


def oracle_cfg(oracle, k, **cfg):
    """Translate gs_match_cfg style keywords to the oracle's MatchConfig."""
    return oracle.match_cfg(
        k=k, classify=bool(cfg.get("classify_reads", 1)), max_paths=cfg.get("max_classification_paths", 10),
        max_read_tax_err=cfg.get("max_read_tax_error_count", -1.0), max_read_class_err=cfg.get("max_read_class_error_count", -1.0),
        threshold=cfg.get("min_kmers_for_class", 1), max_kmer_res_counts=cfg.get("max_kmer_res_counts", 0),
        count_unique=bool(cfg.get("count_unique_kmers", 1)), write_kraken=bool(cfg.get("want_runs", 0)), write_filtered=True,
        use_filter=bool(cfg.get("use_bloom_filter", 1)), dump_labels=cfg.get("dump_labels", False))


def assert_match_parity(native, orun, res, counts, top=None, check_unique=True):
    """Bit-exact comparison of the integer outputs of a GPU match run with an oracle Run."""
    n = orun.n_reads
    assert len(res) == n
    o = orun.reads
    np.testing.assert_array_equal(res["class_vidx"], o["class_vidx"], err_msg="class node")
    np.testing.assert_array_equal(res["tax_err"].astype(np.int64), np.where(o["tax_err"] < 0, 0xFFFFFFFF, o["tax_err"].astype(np.int64)), err_msg="readTaxErrorCount")
    acc = (res["flags"] & native.GS_READ_ACCEPTED) != 0
    np.testing.assert_array_equal(acc, o["accepted"] != 0, err_msg="accepted")
    np.testing.assert_array_equal(res["read_kmers"][acc].astype(np.int64), o["read_kmers"][o["accepted"] != 0].astype(np.int64), err_msg="readKmers")
    s = orun.stats
    for name, field in (("kmers", "kmers"), ("contigs", "contigs"), ("sqsum", "contig_len_squared_sum"), ("reads1", "reads_1kmer"),
                        ("reads", "reads"), ("reads_kmers", "reads_kmers"), ("reads_bps", "reads_bps")):
        np.testing.assert_array_equal(counts[field], s[name], err_msg=name)
    np.testing.assert_array_equal(counts["max_contig_len"].astype(np.int64), s["maxlen"], err_msg="maxContigLen")
    np.testing.assert_array_equal(counts["touched"], orun.has_stats, err_msg="statsIndex != null")
    if check_unique:
        np.testing.assert_array_equal(counts["unique_kmers"], s["unique"], err_msg="unique k-mers")
    if top is not None:
        assert orun.max_counts_n == top.shape[1]
        for v in range(top.shape[0]):
            if orun.max_counts_has[v]:
                np.testing.assert_array_equal(top[v], orun.max_counts[v], err_msg="max kmer counts row %d" % v)
            else:
                assert not top[v].any()


def found_from_filtered(orun_filtered, n):
    """Which reads the oracle wrote to the filtered FASTQ (header `@r<i> ...`) -> bool[n]."""
    found = np.zeros(n, dtype=bool)
    for line in orun_filtered.split(b"\n")[0::4]:
        if line.startswith(b"@r"):
            found[int(line[2:].split(b" ")[0])] = True
    return found


def kraken_from_runs(taxids, run_off, runs, i):
    """Render read i's contig runs the way printKrakenStyleOut does (C/match/FastqKMerMatcher.java:597-611)."""
    parts = []
    for j in range(int(run_off[i]), int(run_off[i + 1])):
        lab, ln = int(runs["label"][j]), int(runs["len"][j])
        name = "A" if lab == 0xFFFFFFFD else ("0" if lab == 0xFFFFFFFE else taxids[lab])
        parts.append("%s:%d" % (name, ln))
    return " ".join(parts)


def small_project(genome_len=20000, seed=7):
    return synth.tiny_project(genome_len=genome_len, seed=seed, shared_frac=0.02)


def host_meta(host, odb):
    """DbMeta of the C++ host layer from the oracle database's metadata (names, ranks, tree, db k-mer counts)."""
    parent, depth, position, has_node = odb.tree()
    return host.DbMeta(odb.k, odb.n_kmers, odb.taxids(), odb.node_names(), odb.node_ranks(), parent, position, depth, has_node, odb.db_kmers())


def bgzf_bytes(data, block=0xff00, level=6, eof_block=True):
    """`data` as block gzip (BGZF, what htslib's bgzip writes): gzip members of at most `block` input bytes, each with the
    'BC' extra subfield that holds the member's size - 1, and the empty end-of-file block."""
    import struct
    import zlib
    out = bytearray()
    pieces = [data[i:i + block] for i in range(0, len(data), block)] + ([b""] if eof_block else [])
    for piece in pieces:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = co.compress(piece) + co.flush()
        total = 18 + len(body) + 8
        assert total <= 0x10000
        out += struct.pack("<4BI2BH2BHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord("B"), ord("C"), 2, total - 1)
        out += body + struct.pack("<II", zlib.crc32(piece) & 0xffffffff, len(piece))
    return bytes(out)
