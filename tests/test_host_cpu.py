"""CPU-side checks of the C++ host layer that need no GPU: FASTQ/FASTA parsing against the oracle, Double.toString."""
import math

import numpy as np
import pytest


def test_java_double_to_string_matches_oracle_and_known_values(native, oracle):
    from genestrip_b200 import host
    cases = {1.0: "1.0", 0.5: "0.5", 100.0: "100.0", 1234567.0: "1234567.0", 1.0e7: "1.0E7", 1.0e-3: "0.001", 1.0e-4: "1.0E-4",
             0.1: "0.1", 1 / 3: "0.3333333333333333", 123456789.125: "1.23456789125E8", 2.0e-5: "2.0E-5", 150.0: "150.0",
             -2.5: "-2.5", 1.7976931348623157e308: "1.7976931348623157E308"}
    for v, s in cases.items():
        assert host.java_double_to_string(v) == s
    import random
    rng = random.Random(1)
    for _ in range(2000):
        v = rng.random() * 10 ** rng.randint(-8, 12)
        assert host.java_double_to_string(v) == oracle.java_double_to_string(v)
        assert float(host.java_double_to_string(v).replace("E", "e")) == v
    assert host.java_double_to_string(math.nan) == "NaN" and host.java_double_to_string(math.inf) == "Infinity"


def _oracle_rewrite(oracle, k, files, is_fasta, with_probs):
    flt = oracle.Bloom(kind=0)   # empty filter: nothing is accepted, every record goes to `rest`
    flt.ensure(10)
    run = oracle.filter_files(flt, k, files, with_probs=with_probs, is_fasta=is_fasta, initial_read_size=256)  # the smallest legal initialReadSizeBytes (C/GSConfigKey.java:348)
    flt.free()
    return run


def test_parser_matches_oracle_on_fixtures_and_quirks(native, oracle, tmp_path):
    import gzip
    import os
    import numpy as np
    from genestrip_b200 import host, synth
    here = os.path.dirname(os.path.abspath(__file__))
    simple = open(os.path.join(here, "golden", "SimpleTest.fastq"), "rb").read()
    dengue_fa = open(os.path.join(here, "golden", "dengue1.fasta"), "rb").read()
    rng = np.random.default_rng(0)
    g = synth.random_genome(rng, 5000).tobytes()
    crlf = b"@a 1\r\n" + g[:150] + b"\r\n+\r\n" + b"I" * 150 + b"\r\n@b\r\n" + g[200:260] + b"\r\n+\r\n" + b"J" * 60 + b"\r\n"
    multi = b"@m desc\n" + g[300:360] + b"\n" + g[360:400] + b"\n\n" + g[400:410] + b"\n+m desc\n" + b"K" * 50 + b"\n" + b"K" * 60 + b"\n"
    nul = b"@n\0x\n" + g[500:540] + b"\0\0" + g[540:600] + b"\n+\n" + b"I" * 100 + b"\n"
    no_final_newline = b"@e\n" + g[700:850] + b"\n+\n" + b"I" * 150
    plus_in_quality = b"@p\n" + g[900:1000] + b"\n+\n" + b"+" + b"I" * 99 + b"\n@q\n" + g[1000:1100] + b"\n+\n" + b"@" * 100 + b"\n"
    fasta = b">f1 x\n" + g[1200:1270] + b"\n" + g[1270:1300] + b"\n\n>f2\n" + g[1400:1500] + b"\n>f3 empty\n>f4\n" + g[1600:1650]
    cases = [([simple], [False]), ([crlf], [False]), ([multi, nul], [False, False]), ([no_final_newline], [False]),
             ([plus_in_quality], [False]), ([fasta], [True]), ([dengue_fa], [True]), ([simple, fasta, crlf], [False, True, False]), ([b""], [False])]
    for files, fa in cases:
        for with_probs in (False, True):
            for k in (2, 31):
                o = _oracle_rewrite(oracle, k, files, fa, with_probs)
                h = host.parse_only(k, files, fa, with_probs)
                assert (h.total_reads, h.total_kmers, h.total_bps) == (o.total_reads, o.total_kmers, o.total_bps)
                assert h.rest == o.rest
    # SimpleTest.fastq: the reference test's expectations (T/fastq/FastqReaderTest.java:43-75)
    h = host.parse_only(2, [simple], [False], True)
    lines = h.rest.split(b"\n")
    assert lines[1] == b"GATTTGGGGTTCAAAGCAGTATCGATCAAATAGTAAATCCATTTGTTCAACTCACAGTTT" and lines[5] == b"CGAT" and lines[7] == b"!**>"
    # files (plain and gzip) give the same records as memory; FASTA entries alternate between the two pooled ReadEntry objects
    p1, p2 = str(tmp_path / "x.fastq"), str(tmp_path / "x.fastq.gz")
    big = synth.fastq_bytes(*synth.sample_reads([g], 5000, 150, seed=2)[:2])
    open(p1, "wb").write(big)
    with gzip.open(p2, "wb") as f:
        f.write(big)
    m = host.parse_only(31, [big])
    assert host.parse_only(31, [p1]).rest == m.rest == host.parse_only(31, [p2]).rest and m.total_reads == 5000
    assert list(host.parse_only(31, [fasta], [True]).accept) == [0, 1, 0, 1]


def test_feeder_record_boundary_search(native):
    """Where the GPU FASTQ feeder cuts a text chunk (feedFastqText / lastRecordStart in gs_host.cpp): at a line that starts
    with '@' whose second-next line starts with '+', searched from the end; a quality line that starts with '@' is never
    taken for a header, and everything before the cut consists of whole records."""
    import numpy as np
    from genestrip_b200 import host, synth
    rng = np.random.default_rng(3)
    g = synth.random_genome(rng, 4000).tobytes()
    recs = []
    for i in range(40):
        L = int(rng.integers(1, 120))
        q = bytes(rng.integers(33, 74, size=L).astype(np.uint8))
        if i % 3 == 0:
            q = b"@" + q[1:]          # quality starting with '@'
        if i % 5 == 0:
            q = b"+" + q[1:]          # ... or with '+'
        recs.append(b"@r%d d\n%s\n+%s\n%s\n" % (i, g[i * 50:i * 50 + L], b"r%d" % i if i % 2 else b"", q))
    text = b"".join(recs)
    starts = np.cumsum([0] + [len(r) for r in recs])
    for cut_at in list(range(len(recs[0]) + len(recs[1]) + len(recs[2]) + 1, len(text), 37)) + [len(text)]:
        c = host.last_record_start(text[:cut_at])
        assert c in starts, "cut %d of %d is not a record start" % (c, cut_at)
        assert c <= cut_at
        # the search finds one of the last records (never further back than the 64 lines it looks at)
        assert cut_at - c <= 64 * 125
    # a complete record must follow the cut's header within the chunk: header + sequence + '+' line seen
    c = host.last_record_start(recs[0] + recs[1][:len(recs[1]) // 4])
    assert c in (0, len(recs[0])) or c == 0
    assert host.last_record_start(b"") == 0
    assert host.last_record_start(b"no newline at all") == 0
    assert host.last_record_start(b"ACGT\nACGT\nACGT\nACGT\n" * 30) == 0     # no '@' header anywhere


def test_block_gzip_reader(native, tmp_path):
    """The feeder inflates block-gzip (BGZF) input with several threads (BgzfReader in gs_host.cpp).  The text
    must be what gzip / java.util.zip.GZIPInputStream give for the same file (a multi-member gzip file,
    C/fastq/AbstractFastqReader.java:224 via StreamProvider), for any request size; bytes behind the last member are
    ignored; an ordinary gzip member ends the fast path exactly there; a corrupt block is an error."""
    import gzip
    import util
    from genestrip_b200 import host
    rng = np.random.default_rng(5)
    recs = []
    for i in range(4000):
        n = int(rng.integers(30, 400))
        seq = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n, p=[0.24, 0.25, 0.25, 0.25, 0.01]).tobytes()
        recs.append(b"@r%d some text\n" % i + seq + b"\n+\n" + bytes(rng.integers(33, 74, size=n, dtype=np.uint8)) + b"\n")
    text = b"".join(recs)
    p = tmp_path / "reads.fastq.gz"
    for block in (0xff00, 7001, 100):
        sub = text if block > 100 else text[:20000]
        p.write_bytes(util.bgzf_bytes(sub, block=block))
        assert gzip.decompress(p.read_bytes()) == sub             # the fixture is an ordinary gzip file
        for request in (1 << 20, 65536, 65537, 4096, 1):
            if request == 1 and len(sub) > 20000:
                continue
            got, foreign = host.bgzf_read_all(p, request)
            assert foreign == -1 and got == sub, (block, request)
    # no end-of-file block, empty blocks in the middle, trailing bytes that are no gzip header
    body = util.bgzf_bytes(text[:50000], eof_block=False) + util.bgzf_bytes(b"", eof_block=True) + util.bgzf_bytes(text[50000:90000], eof_block=False)
    p.write_bytes(body + b"\0\0\0garbage")
    assert host.bgzf_read_all(p, 30000) == (text[:90000], -1)
    # an ordinary gzip member in the middle: the block reader stops exactly in front of it
    head = util.bgzf_bytes(text[:70000], eof_block=False)
    p.write_bytes(head + gzip.compress(text[70000:80000]))
    assert host.bgzf_read_all(p, 1 << 16) == (text[:70000], len(head))
    # corrupt data / truncated file: an error, as from gzread / GZIPInputStream
    bad = bytearray(util.bgzf_bytes(text[:70000]))
    bad[40] ^= 0x55
    p.write_bytes(bytes(bad))
    with pytest.raises(Exception):
        host.bgzf_read_all(p, 1 << 16)
    p.write_bytes(util.bgzf_bytes(text[:70000])[:-40])
    with pytest.raises(Exception):
        host.bgzf_read_all(p, 1 << 16)


def test_device_inflate_decoder_logic_on_the_host(native, tmp_path):
    """The deflate decoder of gs_inflate.cu (block-gzip members inflated on the device), compiled as plain C++ by
    tests/inflate_host_harness.cpp and run thread by thread: stored / fixed / dynamic blocks at every compression level and
    block size against zlib, and every flipped bit detected (size or CRC-32), as gzread / GZIPInputStream would
    (C/fastq/AbstractFastqReader.java:224 reads through them).  The GPU run of the same kernel is tests/test_gpu_inflate.py."""
    import ctypes
    import os
    import subprocess
    import zlib
    import util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "genestrip_b200", "csrc", "gs_inflate.cu")).read()
    body = src.replace('#include "gs_kernels.cuh"', "").replace("#include <cuda_runtime.h>", "")
    body = body[:body.index("void gs_launch_inflate_blocks")]
    inc = tmp_path / "inflate_body.inc"
    inc.write_text(body)
    so = tmp_path / "libinflate_harness.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-shared", "-fPIC",
                    "-DGS_INFLATE_BODY=\"%s\"" % inc, "-o", str(so), os.path.join(root, "tests", "inflate_host_harness.cpp")], check=True)
    L = ctypes.CDLL(str(so))

    def inflate(comp, blocks, n_out):
        comp = np.concatenate([np.frombuffer(comp, dtype=np.uint8), np.zeros(16, dtype=np.uint8)])   # the device buffer has the same slack (aligned 32-bit loads)
        out = np.zeros(max(n_out, 1), dtype=np.uint8)
        blocks = blocks.copy()
        L.gs_inflate_harness_run(comp.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), blocks.ctypes.data_as(ctypes.c_void_p), len(blocks))
        return out[:n_out].tobytes(), blocks

    rng = np.random.default_rng(1)
    recs = []
    for i in range(1500):
        n = int(rng.integers(30, 300))
        seq = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n, p=[0.24, 0.25, 0.25, 0.25, 0.01]).tobytes()
        qual = bytes(rng.integers(33, 74, size=n, dtype=np.uint8)) if i % 3 else b"I" * n
        recs.append(b"@r%d x\n" % i + seq + b"\n+\n" + qual + b"\n")
    fastq = b"".join(recs)
    cases = {"fastq": fastq, "random": bytes(rng.integers(0, 256, size=70000, dtype=np.uint8)), "zeros": b"\0" * 100000, "short": b"A", "empty": b"",
             "text": b"the quick brown fox " * 3000, "skewed": bytes(np.minimum(rng.geometric(0.08, size=60000), 255).astype(np.uint8))}
    for name, data in cases.items():
        for level in (0, 1, 6, 9):
            for block in (0xff00, 1000, 37):
                d = data[:8000] if block < 1000 else data
                comp = util.bgzf_bytes(d, block=block, level=level)
                blocks, n = native.bgzf_blocks(comp)
                out, b = inflate(comp, blocks, n)
                assert n == len(d) and out == d and not b["status"].any(), (name, level, block)
    d = fastq[:60000]
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)      # fixed Huffman codes
    body = co.compress(d) + co.flush()
    out, b = inflate(body, np.array([(0, 0, len(body), len(d), zlib.crc32(d), 0)], dtype=native.DEFLATE_BLOCK_DTYPE), len(d))
    assert out == d and not b["status"].any()
    comp = util.bgzf_bytes(fastq[:150000])
    blocks, n = native.bgzf_blocks(comp)
    for trial in range(120):                                            # any flipped bit is detected
        c = bytearray(comp)
        c[int(rng.integers(18, len(c) - 8 - 28))] ^= 1 << int(rng.integers(0, 8))
        out, b = inflate(bytes(c), blocks, n)
        assert b["status"].any() or out == fastq[:150000]
    for field, delta, want in (("in_len", -5, 1), ("out_len", -5, 2), ("crc32", 1, 3)):
        b2 = blocks.copy()
        b2[field][0] = int(b2[field][0]) + delta
        assert inflate(comp, b2, n)[1]["status"][0] == want


def _pack_reference(b):
    """Plain numpy statement of gs_pack.hpp: C=0 G=1 A=2 T=3 (C/util/CGAT.java:66-69), anything else invalid (:60-69)."""
    n = len(b)
    words = (n + 31) // 32
    code = np.zeros(words * 32, dtype=np.uint64)
    ok = np.zeros(words * 32, dtype=bool)
    for ch, c in ((ord("C"), 0), (ord("G"), 1), (ord("A"), 2), (ord("T"), 3)):
        m = b == ch
        code[:n][m] = c
        ok[:n] |= m
    shifts = (62 - 2 * np.arange(32, dtype=np.uint64)).astype(np.uint64)
    codes = np.bitwise_or.reduce(code.reshape(words, 32) << shifts[None, :], axis=1) if words else np.zeros(0, dtype=np.uint64)
    valid = (ok.reshape(words, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint32) if words else np.zeros(0, dtype=np.uint32)
    return codes.astype(np.uint64), valid


@pytest.mark.parametrize("threads", [1, 3])
def test_host_packer_matches_numpy_statement(native, threads):
    """The 2-bit packer behind gs_match_cfg.host_pack_threads (no GPU needed): every byte value, ragged tails, empty input,
    several pool threads over a batch larger than one claim."""
    rng = np.random.default_rng(7)
    assert native.lib().gs_pack_isa() in (b"avx512", b"avx2", b"scalar")
    cases = [np.zeros(0, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.frombuffer(b"ACGTacgtNn\r\n" * 11, dtype=np.uint8)]
    for n in (1, 31, 32, 33, 63, 64, 65, 1000, 4097, 700_001):
        b = np.frombuffer(b"CGAT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        bad = rng.random(n) < 0.01
        b[bad] = rng.integers(0, 256, int(bad.sum()), dtype=np.uint8)
        cases.append(b)
    for b in cases:
        codes, valid = native.pack_bases(b, threads=threads)
        rc, rv = _pack_reference(b)
        np.testing.assert_array_equal(valid, rv)
        np.testing.assert_array_equal(codes, rc)
    # unaligned source pointer
    b = np.frombuffer(b"CGAT", dtype=np.uint8)[rng.integers(0, 4, 5003)].copy()
    for sh in (1, 7, 13):
        codes, valid = native.pack_bases(b[sh:], threads=threads)
        rc, rv = _pack_reference(b[sh:])
        np.testing.assert_array_equal(codes, rc)
        np.testing.assert_array_equal(valid, rv)
