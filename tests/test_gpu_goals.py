"""GPU parity of the whole goals through the C++ host layer: `match` CSV / filtered FASTQ / kraken-style output and the
`filter` goal's FASTQ outputs must equal the oracle's (reference at threads=0) byte for byte."""
import gzip
import os

import numpy as np
import pytest

from genestrip_b200 import synth

import util
from test_oracle_golden import dengue1_project

pytestmark = pytest.mark.gpu

K = 31


@pytest.fixture(scope="module")
def host(native):
    from genestrip_b200 import host as h
    return h


@pytest.fixture(scope="module")
def project(oracle, native, gpu_ctx, host):
    nodes, names, genomes = util.small_project(genome_len=50000, seed=21)
    leaves = [t for t, _ in genomes]
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, K, nodes, names, genomes, requested=leaves[:3], index_fpp=1e-6)
    meta = util.host_meta(host, odb)
    yield odb, gdb, meta, genomes
    meta.free()
    gdb.close()
    odb.free()


def _reads(genomes, n, seed, **kw):
    bases, offsets, src = synth.sample_reads([g for _, g in genomes], n, 150, seed=seed, n_rate=0.002, **kw)
    return bases, offsets, src


def _assert_csv_equal(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    assert len(la) == len(lb)
    for x, y in zip(la, lb):
        assert x == y


CASES = [
    dict(),
    dict(count_unique_kmers=0),
    dict(classify_reads=0),
    dict(max_read_tax_error_count=0.3, max_read_class_error_count=0.4),
    dict(min_kmers_for_class=4, max_classification_paths=3),
    dict(max_kmer_res_counts=5),
    dict(layout=1),
]


@pytest.mark.parametrize("cfg", CASES, ids=[",".join("%s=%s" % kv for kv in c.items()) or "default" for c in CASES])
def test_match_goal_csv_filtered_kraken(project, oracle, host, cfg):
    odb, gdb, meta, genomes = project
    b1, o1, s1 = _reads(genomes, 3000, 1)
    b2, o2, s2 = _reads(genomes, 1500, 2, len_jitter=60)
    files = [synth.fastq_bytes(b1, o1, s1), synth.fastq_bytes(b2, o2, s2, prefix="q")]
    ocfg = util.oracle_cfg(oracle, K, want_runs=1, **cfg)
    orun = odb.match_files(ocfg, files)
    res = host.match_goal(gdb, meta, files, write_filtered=True, write_kraken=True, batch_reads=700, **cfg)
    assert res.launches > 0
    assert (res.total_reads, res.total_kmers, res.total_bps) == (orun.total_reads, orun.total_kmers, orun.total_bps)
    assert res.filtered == orun.filtered
    assert res.kraken == orun.kraken
    # the four double sums, accumulated on the host in read order, are bit-identical
    for i in range(4):
        np.testing.assert_array_equal(res.dsums[i], orun.dstats[i])
    _assert_csv_equal(res.csv, orun.csv)


def test_match_goal_fasta_gzip_paths_and_probs(project, oracle, host, tmp_path):
    odb, gdb, meta, genomes = project
    b, o, s = _reads(genomes, 800, 3)
    rng = np.random.default_rng(4)
    qual = bytes(rng.integers(33, 74, size=int(o[-1])).astype(np.uint8))
    bb = b.tobytes()
    fq = b"".join(b"@r%d %d\n%s\n+\n%s\n" % (i, s[i], bb[int(o[i]):int(o[i + 1])], qual[int(o[i]):int(o[i + 1])]) for i in range(len(o) - 1))
    # multi-line FASTA with blank lines, lower-case stretches and no trailing newline (loses its last base, like the reference)
    fa = b"".join(b">f%d some text\n%s\n\n%s\n" % (i, bb[int(o[i]):int(o[i]) + 70], bb[int(o[i]) + 70:int(o[i + 1])]) for i in range(200)) + b">last\nACGTACGTTTGACA"
    files = [fq, fa]
    ocfg = oracle.match_cfg(k=K, write_kraken=True, write_filtered=True, with_probs=True)
    orun = odb.match_files(ocfg, files, is_fasta=[False, True])
    res = host.match_goal(gdb, meta, files, is_fasta=[False, True], write_filtered=True, write_kraken=True, with_probs=1, batch_reads=333)
    assert res.filtered == orun.filtered and res.kraken == orun.kraken
    _assert_csv_equal(res.csv, orun.csv)
    # the same through gzip-compressed files and file outputs
    p1, p2 = str(tmp_path / "a.fastq.gz"), str(tmp_path / "b.fasta")
    with gzip.open(p1, "wb") as f:
        f.write(fq)
    open(p2, "wb").write(fa)
    fo, ko = str(tmp_path / "filtered.fastq.gz"), str(tmp_path / "kraken.out")
    res2 = host.match_goal(gdb, meta, [p1, p2], is_fasta=[False, True], write_filtered=True, write_kraken=True, with_probs=1,
                           filtered_path=fo, kraken_path=ko)
    assert gzip.open(fo, "rb").read() == orun.filtered
    assert open(ko, "rb").read() == orun.kraken
    _assert_csv_equal(res2.csv, orun.csv)


def test_match_goal_parser_quirks(project, oracle, host):
    """CRLF input (the '\\r' stays the last base), multi-line FASTQ, NUL bytes, reads shorter than k, missing final newline."""
    odb, gdb, meta, genomes = project
    g = genomes[1][1]
    recs = [b"@a 1\r\n" + g[100:250] + b"\r\n+\r\n" + b"I" * 150 + b"\r\n",
            b"@b\n" + g[300:360] + b"\n" + g[360:420] + b"\n" + g[420:470] + b"\n+b\n" + b"J" * 100 + b"\n" + b"J" * 70 + b"\n",
            b"@c x y\n" + g[500:520] + b"\n+\n" + b"I" * 20 + b"\n",
            b"@d\0\0 z\n" + g[600:700] + b"\0" + g[700:760] + b"\n+\n" + b"I" * 160 + b"\n",
            b"@e\n" + g[800:950] + b"\n+\n" + b"I" * 150]
    fq = b"".join(recs)
    for with_probs in (False, True):
        ocfg = oracle.match_cfg(k=K, write_kraken=True, write_filtered=True, with_probs=with_probs)
        orun = odb.match_files(ocfg, [fq])
        res = host.match_goal(gdb, meta, [fq], write_filtered=True, write_kraken=True, with_probs=int(with_probs), batch_reads=2)
        assert orun.total_reads == 5
        assert (res.total_reads, res.total_kmers, res.total_bps) == (orun.total_reads, orun.total_kmers, orun.total_bps)
        assert res.filtered == orun.filtered
        assert res.kraken == orun.kraken
        _assert_csv_equal(res.csv, orun.csv)


def test_dengue1_golden_through_cuda(oracle, native, gpu_ctx, host):
    """The reference's own end-to-end golden vector R/projects/dengue1/test.out (T/goals/refseq/DBGoalTest.java:127-142),
    reproduced by the CUDA path + C++ host layer byte for byte."""
    nodes, names, fasta, fastq, golden = dengue1_project()
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, K, nodes, names, [("11053", fasta), ("9606", fasta)], requested=["11053"], fill=[True, False])
    meta = util.host_meta(host, odb)
    try:
        res = host.match_goal(gdb, meta, [fastq], write_kraken=True)
        assert res.kraken == golden
    finally:
        meta.free(); gdb.close(); odb.free()


@pytest.mark.parametrize("kind", ["xor", "murmur", "blocked"])
def test_filter_goal(project, oracle, native, gpu_ctx, host, kind):
    """`filter`: FastqBloomFilter with the index filter over the k-mers of the requested taxa (C/goals/refseq/BloomIndexGoal.java:66-111)."""
    odb, gdb, meta, genomes = project
    if kind == "xor":
        flt_o = odb.index_filter()
    else:
        keys, vals = odb.export()
        flt_o = oracle.Bloom(kind=2 if kind == "murmur" else 0, fpp=1e-5)
        sel = keys[::3]
        flt_o.ensure(len(sel))
        flt_o.put(sel)
    okind, p0, p1, factors, words = flt_o.params()
    assert okind == {"blocked": 0, "xor": 1, "murmur": 2}[kind]
    gflt = native.Filter(gpu_ctx, okind, p0, p1, factors, words)
    try:
        # containsLong parity on stored, random and extreme keys
        keys, _ = odb.export()
        rng = np.random.default_rng(8)
        q = np.concatenate([keys[:4000], rng.integers(0, 1 << 62, size=4000, dtype=np.int64),
                            np.array([0, (1 << 62) - 1, -1, -(1 << 63)], dtype=np.int64)])
        if kind == "xor":  # hash == Long.MIN_VALUE (T/bloom/XORKMerBloomFilterTest.java:50-58)
            q = np.concatenate([q, np.array([int(factors[0]) ^ -(1 << 63)], dtype=np.int64)])
        np.testing.assert_array_equal(gflt.contains(q), flt_o.contains(q))
        b, o, s = _reads(genomes, 3000, 5, frac_db=0.3)
        fq = synth.fastq_bytes(b, o, s)
        extra = b"@short\nACGT\n+\nIIII\n@n\n" + b"N" * 150 + b"\n+\n" + b"I" * 150 + b"\n"
        for (mpc, ratio) in ((1, 0.2), (0, 0.2), (0, 0.0), (3, 0.5), (200, 0.1)):
            orun = oracle.filter_files(flt_o, K, [fq + extra], min_pos_count=mpc, pos_ratio=ratio)
            res = host.filter_goal(gflt, K, [fq + extra], min_pos_count=mpc, pos_ratio=ratio, batch_reads=900)
            np.testing.assert_array_equal(res.accept, orun.accept)
            assert res.filtered == orun.filtered and res.rest == orun.rest
            assert (res.total_reads, res.total_kmers, res.total_bps) == (orun.total_reads, orun.total_kmers, orun.total_bps)
    finally:
        gflt.close()
        if kind != "xor":
            flt_o.free()


@pytest.mark.parametrize("kind", ["xor", "murmur", "blocked"])
def test_filter_index_file_round_trip(project, oracle, native, gpu_ctx, host, tmp_path, kind):
    """gs_filter_save_file / gs_filter_load_file: the flat GSF1 file that stands in for KMerProbFilter.save / load
    (`*_index.ser.gz`, C/goals/LoadIndexGoal.java:92-104).  A loaded index answers containsLong and runs the `filter` goal
    exactly like the uploaded one; the file is also written field by field here (what the Java-side GsfExporter writes)
    and loads to the same answers; truncated, foreign and lying files are refused."""
    import struct
    odb, gdb, meta, genomes = project
    if kind == "xor":
        flt_o = odb.index_filter()
    else:
        keys, vals = odb.export()
        flt_o = oracle.Bloom(kind=2 if kind == "murmur" else 0, fpp=1e-5)
        flt_o.ensure(len(keys[::3]))
        flt_o.put(keys[::3])
    okind, p0, p1, factors, words = flt_o.params()
    g1 = native.Filter(gpu_ctx, okind, p0, p1, factors, words)
    path, path2 = str(tmp_path / "index.gsf"), str(tmp_path / "index_by_hand.gsf")
    try:
        g1.save_file(path)
    finally:
        g1.close()
    fac = np.zeros(0, np.int64) if okind == 0 else np.asarray(factors, dtype=np.int64)
    w = np.asarray(words, dtype=np.int64)
    blob = b"GSF1\0\0\0\0" + struct.pack("<iiqqQQ", 1, okind, int(p0), int(p1), len(fac), len(w)) + fac.tobytes() + w.tobytes()
    open(path2, "wb").write(blob)
    assert open(path, "rb").read() == blob
    keys, _ = odb.export()
    rng = np.random.default_rng(21)
    q = np.concatenate([keys[:3000], rng.integers(0, 1 << 62, size=3000, dtype=np.int64), np.array([0, -1, -(1 << 63)], dtype=np.int64)])
    b, o, s_ = _reads(genomes, 2000, 31, frac_db=0.3)
    fq = synth.fastq_bytes(b, o, s_)
    orun = oracle.filter_files(flt_o, K, [fq], min_pos_count=1, pos_ratio=0.2)
    for pth in (path, path2):
        g2 = native.Filter.load_file(gpu_ctx, pth)
        try:
            np.testing.assert_array_equal(g2.contains(q), flt_o.contains(q))
            res = host.filter_goal(g2, K, [fq], min_pos_count=1, pos_ratio=0.2, batch_reads=700)
            np.testing.assert_array_equal(res.accept, orun.accept)
            assert res.filtered == orun.filtered and res.rest == orun.rest
        finally:
            g2.close()
    bad = str(tmp_path / "bad.gsf")
    for data in (blob[:100], b"not an index", blob[:8] + struct.pack("<iiqqQQ", 1, okind, int(p0), int(p1), len(fac), len(w) + (1 << 40)) + blob[48:],
                 blob[:8] + struct.pack("<iiqqQQ", 1, 7, int(p0), int(p1), len(fac), len(w)) + blob[48:]):
        open(bad, "wb").write(data)
        with pytest.raises(native.GenestripError):
            native.Filter.load_file(gpu_ctx, bad)
    if kind != "xor":
        flt_o.free()


@pytest.mark.parametrize("with_probs", [False, True], ids=["noprobs", "probs"])
def test_match_goal_gpu_fastq_feeder(project, oracle, host, tmp_path, with_probs):
    """`match` with the GPU FASTQ feeder (text chunks split on the device) == the oracle (sequential parser + matchRead):
    CSV, filtered FASTQ, totals and the four double sums; many small chunks, two inputs, one of them gzip on disk."""
    odb, gdb, meta, genomes = project
    b1, o1, s1 = _reads(genomes, 4000, 11)
    b2, o2, s2 = _reads(genomes, 1500, 12, len_jitter=60)
    rng = np.random.default_rng(6)
    qual = bytes(rng.integers(33, 74, size=int(o1[-1])).astype(np.uint8))
    bb = b1.tobytes()
    f1 = b"".join(b"@r%d %d\n%s\n+\n%s\n" % (i, s1[i], bb[int(o1[i]):int(o1[i + 1])], qual[int(o1[i]):int(o1[i + 1])]) for i in range(len(o1) - 1))
    f2 = synth.fastq_bytes(b2, o2, s2, prefix="q", qual=b"@")   # quality lines that start with '@' (record-boundary ambiguity)
    p2 = str(tmp_path / "b.fastq.gz")
    with gzip.open(p2, "wb") as f:
        f.write(f2)
    ocfg = oracle.match_cfg(k=K, write_filtered=True, write_kraken=True, with_probs=with_probs)
    orun = odb.match_files(ocfg, [f1, f2])
    p1 = str(tmp_path / "a.fastq")   # plain file: parallel pread() straight into the pinned chunks
    open(p1, "wb").write(f1)
    for chunk in (30000, 1 << 20):
        res = host.match_goal(gdb, meta, [f1 if chunk == 30000 else p1, p2], write_filtered=True, write_kraken=True, with_probs=int(with_probs),
                              text_chunk_bytes=chunk)
        assert res.text_chunks_refused == 0 and res.text_chunks >= (40 if chunk == 30000 else 2)
        assert (res.total_reads, res.total_kmers, res.total_bps) == (orun.total_reads, orun.total_kmers, orun.total_bps)
        assert res.filtered == orun.filtered
        assert res.kraken == orun.kraken   # kraken-style lines from the contig runs of text batches
        for i in range(4):
            np.testing.assert_array_equal(res.dsums[i], orun.dstats[i])
        _assert_csv_equal(res.csv, orun.csv)
    # host-only parsing gives the same bytes
    res0 = host.match_goal(gdb, meta, [f1, p2], write_filtered=True, with_probs=int(with_probs), gpu_parse=False)
    assert res0.text_chunks == 0 and res0.filtered == orun.filtered
    _assert_csv_equal(res0.csv, orun.csv)


def test_match_goal_gpu_fastq_feeder_falls_back(project, oracle, host, tmp_path):
    """Inputs that are not strict 4-line FASTQ: the device refuses the chunk and the sequential parser takes over from
    exactly there -- a strict prefix stays on the GPU, the results are those of the reference parser throughout."""
    odb, gdb, meta, genomes = project
    g = genomes[1][1]
    b1, o1, s1 = _reads(genomes, 1200, 13)
    strict = synth.fastq_bytes(b1, o1, s1)
    odd = [b"@a 1\r\n" + g[100:250] + b"\r\n+\r\n" + b"I" * 150 + b"\r\n",
           b"@b\n" + g[300:360] + b"\n" + g[360:420] + b"\n" + g[420:470] + b"\n+b\n" + b"J" * 100 + b"\n" + b"J" * 70 + b"\n",
           b"@c x y\n" + g[500:520] + b"\n+\n" + b"I" * 20 + b"\n",
           b"@d\0\0 z\n" + g[600:700] + b"\0" + g[700:760] + b"\n+\n" + b"I" * 160 + b"\n"]
    tail = b"@e\n" + g[800:950] + b"\n+\n" + b"I" * 150   # no final newline: the reference drops the last quality byte
    for name, fq, min_gpu in (("odd records in the middle", strict + b"".join(odd) + strict + tail, 3),
                              ("only the tail", strict + tail, 3),
                              ("multi-line from the start", odd[1] + strict, 0)):
        ocfg = oracle.match_cfg(k=K, write_filtered=True, with_probs=True)
        orun = odb.match_files(ocfg, [fq])
        path = str(tmp_path / "in.fastq")
        open(path, "wb").write(fq)
        for src in (fq, path):   # memory, and a plain file (the sequential parser re-opens it behind the consumed bytes)
            res = host.match_goal(gdb, meta, [src], write_filtered=True, with_probs=1, text_chunk_bytes=50000, batch_reads=500)
            assert res.text_chunks_refused == 1, name
            assert res.filtered == orun.filtered, name
            _assert_csv_equal(res.csv, orun.csv)
        assert res.text_chunks_refused == 1, name
        assert res.text_chunks - res.text_chunks_refused >= min_gpu, name
        assert (res.total_reads, res.total_kmers, res.total_bps) == (orun.total_reads, orun.total_kmers, orun.total_bps), name
        assert res.filtered == orun.filtered, name
        _assert_csv_equal(res.csv, orun.csv)


def test_match_goal_block_gzip_input(project, oracle, host, tmp_path):
    """Block-gzip (BGZF) FASTQ files: the feeder inflates the blocks with several host threads (BgzfReader, gs_host.cpp) and
    the records are split on the GPU.  Same results as the oracle's sequential reader over the same text -- also when an
    ordinary gzip member follows the blocks (zlib takes over there) and when a chunk is refused in the middle of the file
    (the sequential parser continues behind the inflated bytes)."""
    import util
    odb, gdb, meta, genomes = project
    g = genomes[1][1]
    b1, o1, s1 = _reads(genomes, 6000, 21)
    b2, o2, s2 = _reads(genomes, 900, 22, len_jitter=40)
    t1, t2 = synth.fastq_bytes(b1, o1, s1), synth.fastq_bytes(b2, o2, s2, prefix="q")
    odd = b"@b\n" + g[300:360] + b"\n" + g[360:420] + b"\n+b\n" + b"J" * 100 + b"\n" + b"J" * 20 + b"\n"   # multi-line record
    cases = (("blocks only", t1 + t2, util.bgzf_bytes(t1 + t2), 0),
             ("small blocks, no end-of-file block", t1, util.bgzf_bytes(t1, block=5000, eof_block=False), 0),
             ("blocks, then an ordinary gzip member", t1 + t2, util.bgzf_bytes(t1, eof_block=False) + gzip.compress(t2), 0),
             ("refused chunk in the middle", t1 + odd + t2, util.bgzf_bytes(t1 + odd + t2, block=20000), 1))
    for name, text, packed, refused in cases:
        assert gzip.decompress(packed) == text
        orun = odb.match_files(oracle.match_cfg(k=K, write_filtered=True, write_kraken=True, with_probs=True), [text])
        path = str(tmp_path / "in.fastq.gz")
        open(path, "wb").write(packed)
        for chunk in (100000, 1 << 22):
            res = host.match_goal(gdb, meta, [path], write_filtered=True, write_kraken=True, with_probs=1, text_chunk_bytes=chunk)
            assert res.text_chunks_refused == refused, name
            assert res.text_chunks - res.text_chunks_refused >= (1 if chunk == 100000 or not refused else 0), name   # (one big chunk: all of it refused)
            assert (res.total_reads, res.total_kmers, res.total_bps) == (orun.total_reads, orun.total_kmers, orun.total_bps), name
            assert res.filtered == orun.filtered, name
            assert res.kraken == orun.kraken, name
            _assert_csv_equal(res.csv, orun.csv)


def test_feeder_inflates_on_the_device(project, oracle, host, tmp_path, monkeypatch):
    """The `match` goal on a block-gzip file: the feeder hands the blocks of every chunk to the device and gets the text
    back; results identical with the host-thread inflater and with the oracle."""
    odb, gdb, meta, genomes = project
    bases, offsets, src = synth.sample_reads([g for _, g in genomes], 40000, 150, seed=23, n_rate=0.002)
    text = synth.fastq_bytes(bases, offsets, src)
    orun = odb.match_files(oracle.match_cfg(k=K, write_filtered=True, write_kraken=True), [text])
    path = str(tmp_path / "in.fastq.gz")
    open(path, "wb").write(util.bgzf_bytes(text, block=20000, level=1))
    before = host.device_inflated_blocks()
    res = host.match_goal(gdb, meta, [path], write_filtered=True, write_kraken=True, text_chunk_bytes=2 << 20)
    assert host.device_inflated_blocks() - before >= len(text) // 20000 - 64     # all but the chunk-boundary blocks
    assert res.text_chunks >= 5 and res.text_chunks_refused == 0
    monkeypatch.setenv("GS_GPU_INFLATE", "0")
    before = host.device_inflated_blocks()
    res_host = host.match_goal(gdb, meta, [path], write_filtered=True, write_kraken=True, text_chunk_bytes=2 << 20)
    assert host.device_inflated_blocks() == before
    for r in (res, res_host):
        assert (r.total_reads, r.total_kmers, r.total_bps) == (orun.total_reads, orun.total_kmers, orun.total_bps)
        assert r.filtered == orun.filtered and r.kraken == orun.kraken
    assert res.csv == res_host.csv


def test_filter_goal_gpu_fastq_feeder(project, oracle, native, gpu_ctx, host):
    """`filter` with the GPU FASTQ feeder: accepted / rejected FASTQ byte-identical with the oracle (ReadEntry.write,
    C/fastq/AbstractFastqReader.java:570-584), with and without qualities, incl. the fall-back on a non-strict tail."""
    odb, gdb, meta, genomes = project
    flt_o = odb.index_filter()
    okind, p0, p1, factors, words = flt_o.params()
    gflt = native.Filter(gpu_ctx, okind, p0, p1, factors, words)
    try:
        b, o, s = _reads(genomes, 4000, 15, frac_db=0.3)
        rng = np.random.default_rng(9)
        qual = bytes(rng.integers(33, 74, size=int(o[-1])).astype(np.uint8))
        bb = b.tobytes()
        fq = b"".join(b"@r%d %d\n%s\n+r%d\n%s\n" % (i, s[i], bb[int(o[i]):int(o[i + 1])], i, qual[int(o[i]):int(o[i + 1])]) for i in range(len(o) - 1))
        tail = b"@last\n" + genomes[0][1][100:250] + b"\n+\n" + b"I" * 150   # no trailing newline
        for text, refused in ((fq, 0), (fq + tail, 1)):
            for with_probs in (False, True):
                orun = oracle.filter_files(flt_o, K, [text], min_pos_count=1, pos_ratio=0.2, with_probs=with_probs)
                res = host.filter_goal(gflt, K, [text], with_probs=with_probs, text_chunk_bytes=60000, batch_reads=900)
                assert res.text_chunks >= 10 and res.text_chunks_refused == refused
                np.testing.assert_array_equal(res.accept, orun.accept)
                assert res.filtered == orun.filtered and res.rest == orun.rest
                assert (res.total_reads, res.total_kmers, res.total_bps) == (orun.total_reads, orun.total_kmers, orun.total_bps)
        # the same text as a block-gzip file: inflated on the device the filter index lives on (gs_filter_context)
        import tempfile
        orun = oracle.filter_files(flt_o, K, [fq + tail], min_pos_count=1, pos_ratio=0.2, with_probs=True)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "in.fastq.gz")
            open(path, "wb").write(util.bgzf_bytes(fq + tail, block=4000))
            before = host.device_inflated_blocks()
            res = host.filter_goal(gflt, K, [path], with_probs=True, text_chunk_bytes=400000, batch_reads=900)
            assert host.device_inflated_blocks() - before >= 64
        assert res.text_chunks_refused == 1
        np.testing.assert_array_equal(res.accept, orun.accept)
        assert res.filtered == orun.filtered and res.rest == orun.rest
    finally:
        gflt.close()


def test_fastq_feeder_fuzz_against_reference_parser(project, oracle, host):
    """Random FASTQ-like inputs -- strict records mixed with CRLF, multi-line sequences and qualities, empty lines, '@' / '+'
    at the start of quality lines, NUL bytes, truncated tails -- at random chunk sizes: whatever the device accepts or
    refuses, the goal's outputs are those of the sequential reference parser + matchRead."""
    odb, gdb, meta, genomes = project
    rng = np.random.default_rng(2024)
    g = genomes[2][1]
    ocfg = oracle.match_cfg(k=K, write_filtered=True, with_probs=True)

    def record(i):
        L = int(rng.integers(0, 220))
        a = int(rng.integers(0, len(g) - 300))
        seq = bytearray(g[a:a + L])
        kind = int(rng.integers(0, 20))
        eol = b"\r\n" if kind == 1 else b"\n"
        qual = bytes(rng.integers(33, 74, size=L).astype(np.uint8))
        if L and kind == 2:
            qual = b"@" + qual[1:]
        if L and kind == 3:
            qual = b"+" + qual[1:]
        hdr = b"@f%d %d" % (i, kind)
        plus = b"+" + (hdr[1:] if kind == 4 else b"")
        if kind == 5 and L > 40:      # multi-line sequence and quality
            return hdr + b"\n" + bytes(seq[:30]) + b"\n" + bytes(seq[30:]) + b"\n+\n" + qual[:50] + b"\n" + qual[50:] + b"\n"
        if kind == 6 and L > 10:      # NUL bytes in the sequence
            seq[5] = 0
        if kind == 7:                 # quality longer than the sequence
            qual = qual + b"II"
        if kind == 8 and L > 10:      # lower case / N
            seq[3] = ord("n"); seq[7] = ord("N")
        return hdr + eol + bytes(seq) + eol + plus + eol + qual + eol

    n_refused = n_gpu = 0
    for trial in range(80):
        n = int(rng.integers(1, 400))
        strict_only = trial % 2 == 0
        recs = []
        for i in range(n):
            r = record(i)
            if strict_only and (b"\0" in r or r.count(b"\n") != 4):
                continue
            recs.append(r)
        text = b"".join(recs)
        if trial % 5 == 1 and text:
            # truncated tail, but only inside the last quality line: a file that ends inside a header / sequence / '+' line (or
            # with a stray empty line) makes the reference index read[-1] (AbstractFastqReader.java:295-307) -- outside the contract
            last_line = len(text) - (text.rfind(b"\n", 0, len(text) - 1) + 1)
            text = text[:-int(rng.integers(1, last_line + 1))]
        chunk = int(rng.integers(4096, 60000))
        orun = odb.match_files(ocfg, [text])
        res = host.match_goal(gdb, meta, [text], write_filtered=True, with_probs=1, text_chunk_bytes=chunk, batch_reads=97)
        assert (res.total_reads, res.total_kmers, res.total_bps) == (orun.total_reads, orun.total_kmers, orun.total_bps), trial
        assert res.filtered == orun.filtered, trial
        _assert_csv_equal(res.csv, orun.csv)
        n_refused += res.text_chunks_refused
        n_gpu += res.text_chunks - res.text_chunks_refused
    assert n_gpu > 20 and n_refused > 5   # both paths were exercised


def test_match_goal_on_the_reference_sample_reads(oracle, native, gpu_ctx, host, tmp_path):
    """Real data: the reference's own sample (data/projects/human_virus/fastq/sample.fastq.gz -- 6565 Illumina reads of variable
    length with real headers, qualities and N runs) through the whole `match` goal: as the gzip file it ships as (zlib's
    sequential reader), re-packed as block gzip (device inflater + GPU record splitter) and as plain text, against the oracle on
    a database built from a third of those reads.  The totals row must be the one the reference's README prints
    (README.md:169: 6565 reads, 658 255 bps, 461 305 k-mers)."""
    here = os.path.dirname(os.path.abspath(__file__))
    gz_path = os.path.join(here, "golden", "human_virus_sample.fastq.gz")
    text = gzip.decompress(open(gz_path, "rb").read())
    lines = text.split(b"\n")
    seqs = [lines[i] for i in range(1, len(lines), 4) if i < len(lines) - 1]
    assert len(seqs) == 6565
    nodes, names, genomes = util.small_project(genome_len=1000, seed=3)
    leaves = [t for t, _ in genomes]
    # every third read (upper-case letters only survive the store's k-mer window) becomes "genome" text of one of the leaf taxa
    parts = {t: [] for t in leaves}
    for i in range(0, len(seqs), 3):
        parts[leaves[(i // 3) % len(leaves)]].append(seqs[i])
    genomes = [(t, b"N".join(parts[t])) for t in leaves]
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, K, nodes, names, genomes)
    meta = util.host_meta(host, odb)
    try:
        ocfg = oracle.match_cfg(k=K, write_filtered=True, write_kraken=True, with_probs=True)
        orun = odb.match_files(ocfg, [text])
        assert (orun.total_reads, orun.total_bps, orun.total_kmers) == (6565, 658255, 461305)
        plain = str(tmp_path / "sample.fastq")
        open(plain, "wb").write(text)
        bgzf = str(tmp_path / "sample.bgzf.fastq.gz")
        open(bgzf, "wb").write(util.bgzf_bytes(text, block=30000))
        hit_rows = 0
        for src, chunk in ((gz_path, 1 << 20), (bgzf, 200000), (plain, 150000), (text, 1 << 22)):
            res = host.match_goal(gdb, meta, [src], write_filtered=True, write_kraken=True, with_probs=1, text_chunk_bytes=chunk)
            assert (res.total_reads, res.total_bps, res.total_kmers) == (6565, 658255, 461305)
            assert res.filtered == orun.filtered
            assert res.kraken == orun.kraken
            for i in range(4):
                np.testing.assert_array_equal(res.dsums[i], orun.dstats[i])
            _assert_csv_equal(res.csv, orun.csv)
            hit_rows = len(res.csv.split(b"\n"))
        assert hit_rows > 5 and len(orun.filtered) > 100000   # a third of the reads are in the database: plenty of hits
        # and with the sequential host parser instead of the GPU feeder
        res0 = host.match_goal(gdb, meta, [gz_path], write_filtered=True, write_kraken=True, with_probs=1, gpu_parse=False)
        assert res0.filtered == orun.filtered and res0.kraken == orun.kraken
        _assert_csv_equal(res0.csv, orun.csv)
    finally:
        meta.free()
        gdb.close()
        odb.free()
