"""GPU parity tests of the `match` path: CUDA (through the C ABI) vs the CPU oracle on the same seeded inputs.

Everything is compared bit-exactly (integer / index work).  The oracle restates
C/match/FastqKMerMatcher.java:327-535 and friends; see oracle/gs_oracle.hpp.
"""
import numpy as np
import pytest

from genestrip_b200 import synth

import util

pytestmark = pytest.mark.gpu

K = 31


@pytest.fixture(scope="module")
def project(oracle, native, gpu_ctx):
    nodes, names, genomes = util.small_project(genome_len=60000, seed=11)
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, K, nodes, names, genomes)
    yield odb, gdb, genomes
    gdb.close()
    odb.free()


@pytest.fixture(scope="module")
def reads(project):
    _, _, genomes = project
    bases, offsets, src = synth.sample_reads([g for _, g in genomes], 6000, 150, seed=4242, frac_db=0.7, sub_rate=0.01, n_rate=0.002)
    return bases, offsets, src, synth.fastq_bytes(bases, offsets, src)


def test_lookup_matches_oracle(project, oracle):
    odb, gdb, genomes = project
    keys, vals = odb.export()
    rng = np.random.default_rng(5)
    hit = keys[rng.integers(0, len(keys), size=5000)]
    miss = rng.integers(0, 1 << 62, size=5000, dtype=np.int64)
    q = np.concatenate([hit, miss, keys[:3], keys[-3:], np.array([0, (1 << 62) - 1], dtype=np.int64)])
    for use_bloom in (True, False):
        odb_lib = oracle.lib()
        odb_lib.gso_db_set_use_filter(odb.h, int(use_bloom))
        v, p = gdb.lookup(q, use_bloom=use_bloom)
        exp = [odb.get(int(x)) for x in q]
        np.testing.assert_array_equal(v, np.array([e[0] for e in exp], dtype=np.int32))
        np.testing.assert_array_equal(p[v >= 0], np.array([e[1] for e in exp], dtype=np.int64)[v >= 0])
    oracle.lib().gso_db_set_use_filter(odb.h, 1)
    # storage position == index in the sorted array (KMerSortedArray.getLong :298-349)
    v, p = gdb.lookup(keys[:1000])
    np.testing.assert_array_equal(p, np.arange(1000))


def test_device_built_bloom_is_bit_identical(project, oracle, native, gpu_ctx):
    odb, _, _ = project
    gdb2 = util.upload(oracle, native, gpu_ctx, odb, device_bloom=True)
    try:
        _, seed, buckets, _, words = odb.store_filter().params()
        np.testing.assert_array_equal(gdb2.bloom_words, words)
    finally:
        gdb2.close()


def test_labels_per_position(project, reads, oracle, native):
    import torch
    odb, gdb, _ = project
    bases, offsets, _, fq = reads
    orun = odb.match_files(oracle.match_cfg(k=K, dump_labels=True), [fq])
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    kofs = np.zeros(len(offsets), dtype=np.uint64)
    kofs[1:] = np.cumsum(np.maximum(lens - K + 1, 0)).astype(np.uint64)
    dev = torch.device("cuda:0")
    pad = np.zeros(64, dtype=np.uint8)
    d_bases = torch.from_numpy(np.concatenate([bases, pad])).to(dev)
    d_off = torch.from_numpy(offsets.view(np.int64)).to(dev)
    d_kofs = torch.from_numpy(kofs.view(np.int64)).to(dev)
    total = int(kofs[-1])
    d_lab = torch.full((total,), -9, dtype=torch.int32, device=dev)
    d_pos = torch.full((total,), -9, dtype=torch.int64, device=dev)
    for layout in (native.GS_LAYOUT_TABLE, native.GS_LAYOUT_CLASSIC):
        sess = native.MatchSession(gdb, native.default_match_cfg(layout=layout))
        try:
            sess.dump_labels(d_bases.data_ptr(), d_off.data_ptr(), len(offsets) - 1, d_kofs.data_ptr(), d_lab.data_ptr(), d_pos.data_ptr())
            torch.cuda.synchronize()
        finally:
            sess.close()
        np.testing.assert_array_equal(d_lab.cpu().numpy(), orun.labels)
        pos = d_pos.cpu().numpy()
        if layout == native.GS_LAYOUT_CLASSIC:
            # storage positions of the reference's sorted array (KMerSortedArray.getLong posStore, :298-349)
            np.testing.assert_array_equal(pos, orun.label_pos)
        else:
            # probe-table slot ids: a bijection of the reference positions (same k-mer <=> same slot)
            hit = orun.label_pos >= 0
            assert (pos[~hit] == -1).all()
            pairs = np.unique(np.stack([orun.label_pos[hit], pos[hit]]), axis=1)
            assert len(np.unique(pairs[0])) == pairs.shape[1] == len(np.unique(pairs[1]))


CONFIGS = [
    dict(),
    dict(use_bloom_filter=0),
    dict(classify_reads=0),
    dict(count_unique_kmers=0),
    dict(max_read_tax_error_count=0.2),
    dict(max_read_tax_error_count=3.0),
    dict(max_read_tax_error_count=0.0),
    dict(max_read_class_error_count=0.3),
    dict(max_read_class_error_count=10.0),
    dict(min_kmers_for_class=5),
    dict(min_kmers_for_class=40, max_read_class_error_count=0.5),
    dict(max_classification_paths=1),
    dict(max_classification_paths=2, min_kmers_for_class=3),
    dict(max_kmer_res_counts=4),
    dict(layout=1),
    dict(layout=1, use_bloom_filter=0),
    dict(layout=1, max_kmer_res_counts=4),
    # how the bases cross the link (gs_match_cfg.host_pack_threads; the default -1 packs them with every available CPU)
    dict(host_pack_threads=0),
    dict(host_pack_threads=1),
    dict(host_pack_threads=3, max_kmer_res_counts=4),
    dict(host_pack_threads=0, layout=1),
    # the split of a batch between the two routes (default: adaptive, starts at 70 % packed)
    dict(host_pack_threads=2, host_pack_percent=100),
    dict(host_pack_threads=2, host_pack_percent=50),
    dict(host_pack_threads=2, host_pack_percent=3, max_kmer_res_counts=4),
    dict(host_pack_threads=2, host_pack_percent=0),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[",".join("%s=%s" % kv for kv in c.items()) or "default" for c in CONFIGS])
def test_match_parity(project, reads, oracle, native, cfg):
    odb, gdb, _ = project
    bases, offsets, _, fq = reads
    orun = odb.match_files(util.oracle_cfg(oracle, K, **cfg), [fq])
    res, ev, counts, top, _, launches = util.gpu_match(native, gdb, bases, offsets, batch=2500, **cfg)
    assert launches > 0
    util.assert_match_parity(native, orun, res, counts, top, check_unique=bool(cfg.get("count_unique_kmers", 1)))
    found = (res["flags"] & native.GS_READ_FOUND) != 0
    np.testing.assert_array_equal(found, util.found_from_filtered(orun.filtered, orun.n_reads))
    # totals of AbstractFastqReader (C/fastq/AbstractFastqReader.java:343-349)
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    assert orun.total_reads == len(lens) and orun.total_bps == lens.sum() and orun.total_kmers == np.maximum(lens - K + 1, 0).sum()
    # maxContigDescriptor: the read named by max_contig_read_no is the first one that reached the maximum
    for v in range(odb.n_values):
        if counts["max_contig_len"][v] > 0:
            assert orun.desc[v] == b"r%d" % counts["max_contig_read_no"][v]
    # the events seen batch by batch end at the same (len, read) pairs
    best = {}
    for e in ev:
        key = int(e["vidx"])
        cand = (int(e["contig_len"]), -int(e["read_no"]))
        if key not in best or cand > best[key]:
            best[key] = cand
    for v in range(odb.n_values):
        if counts["max_contig_len"][v] > 0:
            assert best[v] == (int(counts["max_contig_len"][v]), -int(counts["max_contig_read_no"][v]))


def test_kraken_runs(project, reads, oracle, native):
    odb, gdb, _ = project
    bases, offsets, _, fq = reads
    n = 1500
    sub_off = offsets[:n + 1]
    fq_sub = synth.fastq_bytes(bases, sub_off, None)
    orun = odb.match_files(util.oracle_cfg(oracle, K, want_runs=1), [fq_sub])
    res, _, _, _, runs, _ = util.gpu_match(native, gdb, bases, sub_off, batch=n, want_runs=1)
    run_off, run_arr = runs[0]
    taxids = odb.taxids()
    lines = orun.kraken.split(b"\n")[:-1]
    assert len(lines) == n  # writeAll
    for i, line in enumerate(lines):
        cu, name, tax, size, rest = line.decode().split("\t")
        assert rest == util.kraken_from_runs(taxids, run_off, run_arr, i)
        cls = int(res["class_vidx"][i])
        assert cu == ("C" if cls >= 0 else "U")
        assert tax == (taxids[cls] if cls >= 0 else "0")
        assert int(size) == int(offsets[i + 1] - offsets[i])


def _edge_reads(genomes, rng):
    g0 = np.frombuffer(genomes[0][1], dtype=np.uint8)
    g3 = np.frombuffer(genomes[3][1], dtype=np.uint8)
    reads = []
    reads.append(b"")                                        # empty read
    reads.append(b"ACGT")                                    # shorter than k
    reads.append(g0[100:130].tobytes())                      # L = k-1
    reads.append(g0[100:131].tobytes())                      # L = k: one k-mer
    reads.append(g0[200:350].tobytes().lower())              # lower case is invalid (CGAT.java:60-69)
    reads.append(g0[200:350].tobytes() + b"\r")              # CRLF remnant: last window invalid
    r = bytearray(g0[400:600].tobytes()); r[0] = ord("N"); reads.append(bytes(r))                      # bad base first
    r = bytearray(g0[400:600].tobytes()); r[-1] = ord("N"); reads.append(bytes(r))                     # bad base last
    r = bytearray(g0[400:600].tobytes()); r[169] = ord("N"); reads.append(bytes(r))                    # bad base at max-1
    r = bytearray(g0[400:600].tobytes()); r[170] = ord("N"); reads.append(bytes(r))                    # bad base at max
    r = bytearray(g0[400:600].tobytes()); r[169] = ord("N"); r[180] = ord("x"); reads.append(bytes(r))  # max-1 and tail
    r = bytearray(g0[400:600].tobytes()); r[50] = ord("N"); r[51] = ord("N"); r[90] = ord("N"); reads.append(bytes(r))
    reads.append(b"N" * 150)
    reads.append(b"A" * 150)
    reads.append(bytes([200]) + g0[700:800].tobytes())       # byte >= 0x80 (oracle: invalid)
    # long reads crossing tile boundaries (1024 positions per tile), chimeric between two genera
    for L in (1023 + 30, 1024 + 30, 1025 + 30, 2048 + 30, 2049 + 30, 5000, 10000):
        a = g0[1000:1000 + L // 2]
        b = g3[2000:2000 + (L - L // 2)]
        r = bytearray(np.concatenate([a, b]).tobytes())
        for p in rng.integers(0, L, size=max(1, L // 300)):
            r[int(p)] = ord("CGAT"[int(rng.integers(0, 4))])
        reads.append(bytes(r))
        r2 = bytearray(r)
        r2[min(1024 + 29, L - 1)] = ord("N")
        reads.append(bytes(r2))
    return reads


def _pack(reads):
    offsets = np.zeros(len(reads) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in reads]).astype(np.uint64)
    bases = np.frombuffer(b"".join(reads) + b"\0", dtype=np.uint8)[:-1].copy() if reads else np.zeros(0, dtype=np.uint8)
    return bases, offsets


def _fastq(reads):
    return b"".join(b"@r%d x\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))


@pytest.mark.parametrize("cfg", [dict(), dict(max_read_tax_error_count=0.5), dict(want_runs=1), dict(min_kmers_for_class=3), dict(layout=1),
                                 dict(host_pack_threads=0), dict(host_pack_threads=0, want_runs=1), dict(host_pack_percent=100), dict(host_pack_percent=50, want_runs=1)],
                         ids=["default", "taxerr", "runs", "threshold", "classic", "ascii-link", "ascii-link-runs", "packed-link", "split-link-runs"])
def test_edge_case_reads(project, oracle, native, cfg):
    odb, gdb, genomes = project
    rng = np.random.default_rng(99)
    reads = _edge_reads(genomes, rng)
    # the oracle's parser cannot represent an empty sequence line followed by '+': feed reads directly as FASTQ except
    # the empty one, which is checked separately below
    reads_fq = [r for r in reads if len(r) > 0 and b"\r" not in r]
    fq = _fastq(reads_fq)
    bases, offsets = _pack(reads_fq)
    orun = odb.match_files(util.oracle_cfg(oracle, K, **cfg), [fq])
    assert orun.n_reads == len(reads_fq)
    np.testing.assert_array_equal(orun.reads["read_size"], [len(r) for r in reads_fq])
    res, ev, counts, top, runs, _ = util.gpu_match(native, gdb, bases, offsets, batch=7, **cfg)
    util.assert_match_parity(native, orun, res, counts, top)
    if cfg.get("want_runs"):
        taxids = odb.taxids()
        # reads without k-mers print no line until the (single, reused) entry's buffer exists (writeMatchDetails :723-726)
        lines = {l.split(b"\t")[1]: l for l in orun.kraken.split(b"\n")[:-1]}
        i = 0
        for ro, ru in runs:
            for j in range(len(ro) - 1):
                line = lines.get(b"r%d" % i)
                if line is not None:
                    assert line.decode().split("\t")[4] == util.kraken_from_runs(taxids, ro, ru, j), "read %d" % i
                else:
                    assert ro[j + 1] == ro[j]
                i += 1


def test_crlf_and_empty_reads(project, oracle, native):
    """CRLF FASTQ keeps '\\r' as the last base (B/io/BufferedLineReader.java:160-182 splits on '\\n' only)."""
    odb, gdb, genomes = project
    g0 = np.frombuffer(genomes[0][1], dtype=np.uint8)
    seqs = [g0[200:350].tobytes() + b"\r", g0[500:700].tobytes() + b"\r"]
    fq = b"".join(b"@r%d x\r\n%s\n+\r\n%s\r\n" % (i, s, b"I" * (len(s) - 1)) for i, s in enumerate(seqs))
    orun = odb.match_files(util.oracle_cfg(oracle, K), [fq])
    np.testing.assert_array_equal(orun.reads["read_size"], [len(s) for s in seqs])
    bases, offsets = _pack(seqs)
    res, ev, counts, top, _, _ = util.gpu_match(native, gdb, bases, offsets)
    util.assert_match_parity(native, orun, res, counts)
    # empty batch and empty reads through the ABI
    bases, offsets = _pack([b"", b"", g0[200:350].tobytes(), b""])
    res, _, counts, _, _, _ = util.gpu_match(native, gdb, bases, offsets)
    assert (res["flags"][[0, 1, 3]] == 0).all() and (res["class_vidx"][[0, 1, 3]] == -1).all()
    res, _, counts, _, _, _ = util.gpu_match(native, gdb, np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    assert len(res) == 0 and counts["kmers"].sum() == 0


def test_many_taxa_per_read_slow_path(oracle, native, gpu_ctx):
    """> GS_TABLE_CAP (128) distinct taxa in one read: the overflow list + slow path must give identical results."""
    V_LEAVES = 300
    edges = ["1\t|\t1\t|\tno rank\t|\t\t|\n"]
    for g in range(10):
        edges.append("%d\t|\t1\t|\tgenus\t|\t\t|\n" % (100 + g))
    for s in range(V_LEAVES):
        edges.append("%d\t|\t%d\t|\tspecies\t|\t\t|\n" % (1000 + s, 100 + s % 10))
    nodes = "".join(edges)
    names = "".join("%s\t|\tn%s\t|\t\t|\tscientific name\t|\n" % (e.split("\t")[0], e.split("\t")[0]) for e in edges)
    rng = np.random.default_rng(3)
    genomes = [(str(1000 + s), synth.random_genome(rng, 400).tobytes()) for s in range(V_LEAVES)]
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, K, nodes, names, genomes)
    try:
        # chimeric long reads: 40 bases from each of many genomes
        reads = []
        for r in range(6):
            order = rng.permutation(V_LEAVES)[: (50, 129, 200, 300, 128, 299)[r]]
            reads.append(b"".join(genomes[int(g)][1][10:10 + 45] for g in order))
        reads.append(genomes[0][1][:200])
        fq = _fastq(reads)
        bases, offsets = _pack(reads)
        for cfg in (dict(), dict(max_classification_paths=128), dict(min_kmers_for_class=20), dict(max_read_tax_error_count=0.9)):
            orun = odb.match_files(util.oracle_cfg(oracle, K, **cfg), [fq])
            res, ev, counts, top, _, _ = util.gpu_match(native, gdb, bases, offsets, batch=4, **cfg)
            util.assert_match_parity(native, orun, res, counts)
            slow = (res["flags"] & native.GS_READ_SLOWPATH) != 0
            assert slow.sum() >= 3
    finally:
        gdb.close()
        odb.free()


def test_reference_classification_table(oracle, native, gpu_ctx):
    """T/match/FastqKMerMatcherTest.java:315-412 (testReadClassification) through the CUDA path: k=2, tree 1<-2, 1<-3,
    DB {CC->1, CT->2, CG->3}; the expected classes are the reference test's own table."""
    nodes = "".join("%d\t|\t1\t|\tno rank\t|\t\t|\n" % t for t in (1, 2, 3))
    names = "".join("%d\t|\t%d\t|\t\t|\tscientific name\t|\n" % (t, t) for t in (1, 2, 3))
    genomes = [("1", b"CC"), ("2", b"CT"), ("3", b"CG")]
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, 2, nodes, names, genomes)
    taxids = odb.taxids()
    table = [(0, "CCCC", "1"), (0, "GAGAGA", None), (0, "CCCG", "3"), (0, "AGGGG", "2"), (0, "CCCCCCT", "2"),
             (1, "CTCCT", "2"), (1, "CTCTCCT", None), (1, "TAGGGG", "2"), (1, "TAGGGGT", None),
             (0.5, "CCA", "1"), (0.5, "CCAA", None), (0.1, "CC", "1"), (0.1, "CCA", None), (0.1, "CCAA", None),
             (0.99, "TTTT", None), (0.99, "CTTT", "2")]
    try:
        for err in sorted(set(t[0] for t in table)):
            rows = [t for t in table if t[0] == err]
            bases, offsets = _pack([t[1].encode() for t in rows])
            res, _, _, _, _, _ = util.gpu_match(native, gdb, bases, offsets, max_read_tax_error_count=float(err),
                                                max_classification_paths=4, use_bloom_filter=0)
            got = [taxids[c] if c >= 0 else None for c in res["class_vidx"]]
            assert got == [t[2] for t in rows], "error threshold %s" % err
    finally:
        gdb.close()
        odb.free()


def test_reference_match_read_kat(oracle, native, gpu_ctx):
    """T/match/FastqKMerMatcherTest.java:96-210 (testMatchRead): k=2 store {CC->1, TT->2, AG->3}; random 500-bp reads;
    kmers / contigs / maxContigLen / unique per taxon from an independent scan of the read (as the reference test does)."""
    nodes = "".join("%d\t|\t1\t|\tno rank\t|\t\t|\n" % t for t in (1, 2, 3))
    names = "".join("%d\t|\t%d\t|\t\t|\tscientific name\t|\n" % (t, t) for t in (1, 2, 3))
    genomes = [("1", b"CC"), ("2", b"TT"), ("3", b"AG")]
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, 2, nodes, names, genomes)
    taxids = odb.taxids()
    rng = np.random.default_rng(42)
    pairs = {b"CC": "1", b"GG": "1", b"AA": "2", b"TT": "2", b"AG": "3", b"CT": "3"}
    try:
        for _ in range(40):
            read = synth.BASES[rng.integers(0, 4, size=500)].tobytes()
            kmers = {t: 0 for t in "123"}; contigs = dict(kmers); maxlen = dict(kmers)
            last, run = None, 0
            for j in range(499):
                t = pairs.get(read[j:j + 2])
                if t != last:
                    if last is not None:
                        contigs[last] += 1; maxlen[last] = max(maxlen[last], run)
                    run = 0
                if t is not None:
                    kmers[t] += 1; run += 1
                last = t
            if last is not None:
                contigs[last] += 1; maxlen[last] = max(maxlen[last], run)
            bases, offsets = _pack([read])
            _, _, counts, _, _, _ = util.gpu_match(native, gdb, bases, offsets, classify_reads=0, use_bloom_filter=0)
            for v, t in enumerate(taxids):
                assert counts["kmers"][v] == kmers[t] and counts["contigs"][v] == contigs[t] and counts["max_contig_len"][v] == maxlen[t]
                assert counts["unique_kmers"][v] == (1 if kmers[t] else 0)
    finally:
        gdb.close()
        odb.free()


def test_two_sessions_share_one_database(project, reads, oracle, native):
    """The probe table's in-line seen bits are leased by the first unique-counting session; a second concurrent session
    falls back to its own bitset.  Both must give the reference's unique counts, also after the lease is released."""
    odb, gdb, _ = project
    bases, offsets, _, fq = reads
    n = 2000
    off = np.ascontiguousarray(offsets[: n + 1])
    orun = odb.match_files(util.oracle_cfg(oracle, K), [synth.fastq_bytes(bases, off, None)])
    a = native.MatchSession(gdb)
    b = native.MatchSession(gdb)
    try:
        ta = a.submit(bases, off, 0)
        tb = b.submit(bases, off, 0)
        ra, _, _, _ = a.collect(ta)
        rb, _, _, _ = b.collect(tb)
        ca, _ = a.finish()
        cb, _ = b.finish()
    finally:
        a.close()
        b.close()
    util.assert_match_parity(native, orun, ra, ca)
    util.assert_match_parity(native, orun, rb, cb)
    res, _, counts, _, _, _ = util.gpu_match(native, gdb, bases, off)   # new lease: bits were cleared
    util.assert_match_parity(native, orun, res, counts)


@pytest.mark.parametrize("k", [31, 24, 21], ids=["k31", "k24-smallest-with-minimizer-prefilter", "k21-no-prefilter"])
def test_match_parity_other_k(oracle, native, gpu_ctx, k):
    """Lookups, per-taxon counts, unique k-mers and hit counters for other k-mer sizes: the minimizer prefilter needs
    k >= 24 (GS_MZ_MIN_K), below that every k-mer probes the table (KMerSortedArray.getLong :298-349,
    KMerUniqueCounterBits :117-199)."""
    nodes, names, genomes = util.small_project(genome_len=12000, seed=23)
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, k, nodes, names, genomes)
    try:
        keys, _ = odb.export()
        rng = np.random.default_rng(3)
        q = np.concatenate([keys, rng.integers(0, 1 << (2 * k), size=4000, dtype=np.int64)])
        v, p = gdb.lookup(q, use_bloom=False)
        exp = [odb.get(int(x)) for x in q[::7]]
        np.testing.assert_array_equal(v[::7], np.array([e[0] for e in exp], dtype=np.int32))
        assert (p[:len(keys)] == np.arange(len(keys))).all()
        bases, offsets, src = synth.sample_reads([g for _, g in genomes], 3000, 150, seed=77, frac_db=0.8, sub_rate=0.01, n_rate=0.002)
        fq = synth.fastq_bytes(bases, offsets, src)
        for cfg in (dict(), dict(max_kmer_res_counts=4), dict(prefilter=0)):
            ocfg = {kk: vv for kk, vv in cfg.items() if kk != "prefilter"}
            orun = odb.match_files(util.oracle_cfg(oracle, k, **ocfg), [fq])
            res, ev, counts, top, _, _ = util.gpu_match(native, gdb, bases, offsets, batch=1100, **cfg)
            util.assert_match_parity(native, orun, res, counts, top)
    finally:
        gdb.close()
        odb.free()


def test_wide_minimizer_order(oracle, native, gpu_ctx, monkeypatch):
    """Stores beyond ~6e8 k-mers order their minimizers by a 64-bit hash (the 32-bit space is too crowded for a minimum of
    nine); GS_DEBUG_MZ_WIDE forces that mode on a small store: same labels, counts and unique k-mers."""
    nodes, names, genomes = util.small_project(genome_len=15000, seed=29)
    monkeypatch.setenv("GS_DEBUG_MZ_WIDE", "1")
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, K, nodes, names, genomes)
    monkeypatch.delenv("GS_DEBUG_MZ_WIDE")
    try:
        keys, _ = odb.export()
        v, p = gdb.lookup(keys, use_bloom=False)   # includes the self check: every stored key passes the prefilter
        assert (v >= -1).all() and (p == np.arange(len(keys))).all() and (v < 0x7FFFFFF0).all()
        bases, offsets, src = synth.sample_reads([g for _, g in genomes], 4000, 150, seed=78, frac_db=0.6, sub_rate=0.02, n_rate=0.002)
        fq = synth.fastq_bytes(bases, offsets, src)
        for cfg in (dict(), dict(prefilter=0), dict(max_kmer_res_counts=3)):
            ocfg = {kk: vv for kk, vv in cfg.items() if kk != "prefilter"}
            orun = odb.match_files(util.oracle_cfg(oracle, K, **ocfg), [fq])
            res, ev, counts, top, _, _ = util.gpu_match(native, gdb, bases, offsets, batch=1500, **cfg)
            util.assert_match_parity(native, orun, res, counts, top)
    finally:
        gdb.close()
        odb.free()


def test_thread_kernel_hand_over_paths(oracle, native, gpu_ctx):
    """Short-read batches go to the thread-per-read reduce kernel; reads it cannot take -- more than 8 distinct taxa, or
    longer than 2047 bases -- are handed to the warp-per-read kernel, reads with more than 128 taxa on to the slow path.
    A 160-species project with chimeric reads exercises all three in one batch (FastqKMerMatcher.java:327-535)."""
    n_sp = 160
    edges = [(1, 1, "no rank")] + [(10 + g, 1, "genus") for g in range(8)] + [(1000 + s, 10 + s % 8, "species") for s in range(n_sp)]
    nodes = "".join("%d\t|\t%d\t|\t%s\t|\t\t|\n" % e for e in edges)
    names = "".join("%d\t|\ttaxon %d\t|\t\t|\tscientific name\t|\n" % (e[0], e[0]) for e in edges)
    rng = np.random.default_rng(77)
    genomes = [(str(1000 + s), synth.random_genome(rng, 3000).tobytes()) for s in range(n_sp)]
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, K, nodes, names, genomes)
    try:
        reads = []
        for i in range(1500):                      # plain short reads: the thread kernel keeps them
            s = int(rng.integers(0, n_sp)); a = int(rng.integers(0, 2800))
            reads.append(genomes[s][1][a:a + 150])
        for i in range(60):                        # 9 .. 20 taxa in one read (40 bases = 10 k-mers of each): warp kernel
            parts = [genomes[int(s)][1][int(a):int(a) + 40] for s, a in zip(rng.choice(n_sp, size=int(rng.integers(9, 21)), replace=False), rng.integers(0, 2900, size=20))]
            reads.append(b"".join(parts))
        for i in range(120):                       # 2 .. 8 taxa, 3 .. 17 run boundaries, some with an N: both sides of the
            n = 2 + i % 7                          # thread kernel's limit of 16 boundaries per read
            parts = [genomes[int(s)][1][int(a):int(a) + 40] for s, a in zip(rng.choice(n_sp, size=n, replace=False), rng.integers(0, 2900, size=n))]
            r = bytearray(b"".join(parts))
            if i % 3 == 0:
                r[int(rng.integers(0, len(r)))] = ord("N")
            reads.append(bytes(r))
        for i in range(3):                         # more than 128 taxa: slow path (vote table in global memory)
            parts = [genomes[int(s)][1][int(a):int(a) + 32] for s, a in zip(rng.permutation(n_sp)[:150], rng.integers(0, 2900, size=150))]
            reads.append(b"".join(parts))
        for i in range(4):                         # longer than 2047 bases
            s = int(rng.integers(0, n_sp))
            reads.append(genomes[s][1][:2900] if i % 2 else genomes[s][1][100:2148])
        order = rng.permutation(len(reads))
        reads = [reads[int(i)] for i in order]
        bases, offsets = _pack(reads)
        assert len(bases) / len(reads) <= 512      # the batch qualifies for the thread kernel
        fq = _fastq(reads)
        for cfg in (dict(), dict(max_classification_paths=3, min_kmers_for_class=2), dict(max_read_tax_error_count=0.4), dict(max_kmer_res_counts=3)):
            orun = odb.match_files(util.oracle_cfg(oracle, K, **cfg), [fq])
            res, ev, counts, top, _, _ = util.gpu_match(native, gdb, bases, offsets, batch=len(reads), **cfg)
            util.assert_match_parity(native, orun, res, counts, top)
            slow = (res["flags"] & native.GS_READ_SLOWPATH) != 0
            assert slow.sum() == 3
    finally:
        gdb.close()
        odb.free()


def test_radix_upload_gives_same_database(project, reads, oracle, native, gpu_ctx):
    """A RadixKMerStore as the source (C/store/RadixKMerStore.java:369-412, 714-730): buckets by the low r bits, entries =
    (valueIndex << (62 - r)) | (kmer >>> r) sorted by the remaining bits.  gs_db_put_radix_bucket rebuilds the keys and merges
    them into the sorted layout: same positions, values, and match results as the KMerSortedArray upload
    (T/match/RadixKMerStoreBenchmarkTest.java:186-229)."""
    odb, gdb, _ = project
    bases, offsets, _, fq = reads
    keys, vals = odb.export()
    r = 17
    vi = vals.astype(np.int64) + 32768
    bucket = keys & ((1 << r) - 1)
    rem = keys >> r
    order = np.lexsort((rem, bucket))
    b_sorted, entries = bucket[order], (vi[order] << (62 - r)) | rem[order]
    starts = np.flatnonzero(np.concatenate([[True], b_sorted[1:] != b_sorted[:-1]]))
    ends = np.concatenate([starts[1:], [len(b_sorted)]])
    buckets = [(int(b_sorted[a]), np.ascontiguousarray(entries[a:e])) for a, e in zip(starts, ends)]
    parent, depth, position, has_node = odb.tree()
    _, seed, nb, _, words = odb.store_filter().params()
    g2 = native.Database(gpu_ctx, K, None, None, odb.n_values, parent_by_vidx=parent, has_node=has_node, bloom=(seed, nb, words), radix=(r, buckets))
    try:
        assert g2.n_kmers == len(keys)
        np.testing.assert_array_equal(g2.values(), vals)
        v, p = g2.lookup(keys[::11])
        np.testing.assert_array_equal(p, np.arange(len(keys))[::11])
        orun = odb.match_files(util.oracle_cfg(oracle, K), [fq])
        res, ev, counts, top, _, _ = util.gpu_match(native, g2, bases, offsets, batch=3000)
        util.assert_match_parity(native, orun, res, counts, top)
    finally:
        g2.close()


def _unmix62(h):
    """Inverse of gs_mix62 (gs_device.cuh): the bijection on [0, 2^62) that spreads the k-mers over the probe table."""
    M = (1 << 62) - 1
    x = h
    x ^= x >> 33
    x = (x * pow(0x81DADEF4BC2DD44D, -1, 1 << 62)) & M
    x ^= x >> 27
    x ^= x >> 54
    x = (x * pow(0x7FB5D329728EA185, -1, 1 << 62)) & M
    x ^= x >> 31
    return x


def test_probe_table_is_exact_for_displaced_keys(native, gpu_ctx):
    """A key pushed out of its full home bucket sits in the next one.  A lookup whose OWN home is that next bucket and whose
    remainder equals the pushed key's must not take it for its key (a 2^-rbits coincidence at random -- forced here through
    the inverse of the table's hash), and both keys must be found when both are stored."""
    tbits, rbits = 22, 40
    b = 123457
    rems = [5, 77, 900, 12345, 999999, 1 << 39]      # six keys of home bucket b: the last two land in bucket b + 1
    mk = lambda bucket, rem: _unmix62((bucket << rbits) | rem)
    home = [mk(b, r) for r in rems]
    twin_absent = mk(b + 1, rems[-1])                # home b + 1, remainder of a displaced key: NOT stored
    twin_stored = mk(b + 1, rems[-2])                # home b + 1, remainder of the other displaced key: stored too
    far = [mk(b + 1, r) for r in (3, 4)] + [mk(b + 2, r) for r in (8, 9, 10)]   # more pressure: the chain runs on
    stored = sorted(home + [twin_stored] + far)
    keys = np.array(stored, dtype=np.int64)
    vidx = np.arange(len(keys)) % 5
    vals = (vidx - 32768).astype(np.int16)
    parent = np.array([-1, 0, 0, 1, 2], dtype=np.int32)
    db = native.Database(gpu_ctx, K, keys, vals, 5, parent_by_vidx=parent, has_node=None, bloom=None, build_bloom=True)
    try:
        v, p = db.lookup(keys)
        np.testing.assert_array_equal(v, vidx)
        np.testing.assert_array_equal(p, np.arange(len(keys)))
        absent = np.array([twin_absent, mk(b + 2, rems[-1]), mk(b + 3, 9), mk(b, 6)], dtype=np.int64)
        assert not np.isin(absent, keys).any()
        v, _ = db.lookup(absent, use_bloom=False)
        np.testing.assert_array_equal(v, -1)          # 0x7FFFFFFF here = the table disagreed with the sorted array
    finally:
        db.close()
