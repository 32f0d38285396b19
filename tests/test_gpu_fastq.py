"""GPU parity tests of the raw-FASTQ feeder: record splitting on the device (gs_match_submit_fastq) must
give what AbstractFastqReader.doReadFastq (C/fastq/AbstractFastqReader.java:288-368, restated in the oracle) + matchRead
give for the same bytes, and must refuse -- never mis-parse -- anything that is not strict 4-line FASTQ."""
import numpy as np
import pytest

from genestrip_b200 import synth

import util

pytestmark = pytest.mark.gpu

K = 31


@pytest.fixture(scope="module")
def project(oracle, native, gpu_ctx):
    nodes, names, genomes = util.small_project(genome_len=30000, seed=5)
    odb, gdb = util.build_pair(oracle, native, gpu_ctx, K, nodes, names, genomes)
    yield odb, gdb, genomes
    gdb.close()
    odb.free()


def _records(bases, offsets, src, rng, crlf=False, long_qual=False, plus_header=False):
    recs = []
    bb = bases.tobytes()
    eol = b"\r\n" if crlf else b"\n"
    for i in range(len(offsets) - 1):
        seq = bb[int(offsets[i]):int(offsets[i + 1])]
        hdr = b"@r%d %d some text" % (i, int(src[i]))
        qual = b"I" * (len(seq) + (int(rng.integers(0, 4)) if long_qual else 0))
        plus = b"+" + (hdr[1:] if plus_header and i % 3 == 0 else b"")
        recs.append(hdr + eol + seq + eol + plus + eol + qual + eol)
    return recs


def _run_text(native, gdb, chunks, **cfg):
    """Feed text chunks through gs_match_submit_fastq; returns per-read results, record tables, infos, counts."""
    sess = native.MatchSession(gdb, native.default_match_cfg(**cfg))
    res, recs, infos, events = [], [], [], []
    pending = []
    ordinal = 0
    try:
        def collect():
            t, chunk, first = pending.pop(0)
            r, ev, eh, rc = sess.collect_fastq(t)
            res.append(r.copy()); recs.append(rc.copy())
            for e, h in zip(ev, eh):
                line_end = chunk.index(b"\n", int(h))
                events.append((int(e["vidx"]), int(e["contig_len"]), int(e["read_no"]), chunk[int(h):line_end]))
        for chunk in chunks:
            t, info = sess.submit_fastq(np.frombuffer(chunk, dtype=np.uint8), ordinal)
            assert t != 0 and info.status == 0, "chunk refused: status %d" % info.status
            infos.append((info.n_reads, info.total_kmers, info.total_bps))
            pending.append((t, chunk, ordinal))
            ordinal += info.n_reads
            if len(pending) == native.GS_MAX_INFLIGHT:
                collect()
        while pending:
            collect()
        counts, top = sess.finish()
        launches = sess.kernel_launches
    finally:
        sess.close()
    return np.concatenate(res), recs, infos, events, counts, top, launches


@pytest.mark.parametrize("variant", ["plain", "crlf", "long_qual", "plus_header"])
def test_fastq_text_parity(project, oracle, native, variant):
    odb, gdb, genomes = project
    rng = np.random.default_rng(17)
    bases, offsets, src = synth.sample_reads([g for _, g in genomes], 5000, 150, seed=99, frac_db=0.7, sub_rate=0.01, n_rate=0.003)
    recs = _records(bases, offsets, src, rng, crlf=variant == "crlf", long_qual=variant == "long_qual", plus_header=variant == "plus_header")
    # a few odd but strict records: shorter than k, exactly k, lower case, one long read
    g0 = genomes[0][1]
    extra = [b"@short x\nACGT\n+\nIIII\n", b"@exact\n" + g0[100:131] + b"\n+\n" + b"I" * 31 + b"\n",
             b"@lower\n" + g0[200:350].lower() + b"\n+\n" + b"I" * 150 + b"\n", b"@long read\n" + g0[1000:9000] + b"\n+\n" + b"#" * 8000 + b"\n"]
    recs = recs[:2000] + extra + recs[2000:]
    fq = b"".join(recs)
    orun = odb.match_files(util.oracle_cfg(oracle, K), [fq])
    # chunks cut at record boundaries, uneven sizes (one of them a single record)
    cuts = [0, 1, 700, 2003, 2004, 3500, len(recs)]
    chunks = [b"".join(recs[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    res, rtab, infos, events, counts, top, launches = _run_text(native, gdb, chunks)
    assert launches > 0
    assert sum(i[0] for i in infos) == orun.total_reads == len(recs)
    assert sum(i[1] for i in infos) == orun.total_kmers and sum(i[2] for i in infos) == orun.total_bps
    util.assert_match_parity(native, orun, res, counts, top)
    # the record table points at the right bytes
    for chunk, (a, b), tab in zip(chunks, zip(cuts[:-1], cuts[1:]), rtab):
        assert len(tab) == b - a + 1 and int(tab[-1]["hdr_start"]) == len(chunk)
        for j in (0, (b - a) // 2, b - a - 1):
            rec = recs[a + j]
            lines = rec.split(b"\n")
            h, s0, L, q = (int(tab[j][f]) for f in ("hdr_start", "seq_start", "seq_len", "qual_start"))
            assert chunk[h:s0 - 1] == lines[0] and chunk[s0:s0 + L] == lines[1] and chunk[q:int(tab[j + 1]["hdr_start"]) - 1] == lines[3]
    # maxContigDescriptor: the header named by the final (len, lowest read) event of every taxon is the oracle's descriptor
    best = {}
    for v, ln, rn, hdr in events:
        if v not in best or (ln, -rn) > best[v][0]:
            best[v] = ((ln, -rn), hdr)
    for v in range(odb.n_values):
        if counts["max_contig_len"][v] > 0:
            hdr = best[v][1]
            assert hdr.startswith(b"@")
            assert orun.desc[v] == hdr[1:].split(b" ")[0]


def test_fastq_text_contig_runs(project, oracle, native):
    """want_runs through the text path: the per-read contig runs (printKrakenStyleOut, C/match/FastqKMerMatcher.java:597-611)
    equal those of the host-parsed path for the same reads."""
    odb, gdb, genomes = project
    bases, offsets, src = synth.sample_reads([g for _, g in genomes], 1200, 150, seed=31, frac_db=0.8, sub_rate=0.02, n_rate=0.004, len_jitter=80)
    fq = synth.fastq_bytes(bases, offsets, src)
    _, _, _, _, runs_ref, _ = util.gpu_match(native, gdb, bases, offsets, batch=len(offsets) - 1, want_runs=1)
    ro_ref, ru_ref = runs_ref[0]
    sess = native.MatchSession(gdb, native.default_match_cfg(want_runs=1))
    try:
        t, info = sess.submit_fastq(np.frombuffer(fq, dtype=np.uint8))
        assert t and info.status == 0
        res, ev, eh, recs, ro, ru = sess.collect_fastq(t)
        sess.finish()
    finally:
        sess.close()
    np.testing.assert_array_equal(ro, ro_ref)
    np.testing.assert_array_equal(ru, ru_ref)
    assert len(ru) > len(res)


def test_fastq_text_refuses_non_strict_input(project, native):
    _, gdb, genomes = project
    g0 = genomes[0][1]
    ok = b"@a\n" + g0[0:100] + b"\n+\n" + b"I" * 100 + b"\n"
    cases = {
        "multi-line sequence": (b"@a\n" + g0[0:60] + b"\n" + g0[60:100] + b"\n+\n" + b"I" * 100 + b"\n" + ok + ok + ok[:0], native_bits("RECORD", "LINES")),
        "no trailing newline": (ok + ok[:-1], native_bits("TAIL")),
        "NUL byte": (ok + b"@b\nAC\0GT\n+\nIIIII\n", native_bits("NUL")),
        "short quality": (ok + b"@b\nACGTACGT\n+\nIIII\n", native_bits("RECORD")),
        "missing plus": (ok + b"@b\nACGT\nACGT\nIIII\n", native_bits("RECORD")),
        "three lines": (ok + b"@b\nACGT\n+\n", native_bits("LINES")),
        "tiny lines": (b"@\nA\n+\nI\n" * 50, native_bits("CAP")),
    }
    sess = native.MatchSession(gdb, native.default_match_cfg())
    try:
        for name, (text, allowed) in cases.items():
            t, info = sess.submit_fastq(np.frombuffer(text, dtype=np.uint8))
            assert t == 0 and info.status != 0, name
            assert info.status & allowed, "%s: status %d" % (name, info.status)
        # the session is still usable and nothing is pending
        t, info = sess.submit_fastq(np.frombuffer(ok * 5, dtype=np.uint8))
        assert t != 0 and info.n_reads == 5 and info.total_bps == 500 and info.total_kmers == 5 * 70
        res, _, _, tab = sess.collect_fastq(t)
        assert len(res) == 5 and len(tab) == 6
        # empty chunk
        t, info = sess.submit_fastq(np.zeros(0, dtype=np.uint8))
        assert info.status == 0 and info.n_reads == 0
        if t:
            sess.collect_fastq(t)
        sess.finish()
    finally:
        sess.close()


def native_bits(*names):
    bits = {"NUL": 1, "LINES": 2, "CAP": 4, "RECORD": 8, "TAIL": 16}
    v = 0
    for n in names:
        v |= bits[n]
    return v
