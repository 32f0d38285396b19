"""GPU parity of the `db` goal's update phase (C/goals/refseq/DBGoal.java:234-311): starting from the store as the fill phase
leaves it, gs_db_update must produce exactly the values of the oracle's full build (fill + update), and matching against
the updated device database must equal matching against the oracle's."""
import numpy as np
import pytest

from genestrip_b200 import synth

import util

pytestmark = pytest.mark.gpu

K = 31


def _project():
    nodes, names, genomes = util.small_project(genome_len=20000, seed=31)
    rng = np.random.default_rng(12)
    gs = [bytearray(g) for _, g in genomes]
    # lower-case stretches (upper-cased by the store reader, CGAT.java:91-99), N runs and other bytes that reset the window
    for g in gs:
        a = int(rng.integers(0, len(g) - 400))
        g[a:a + 300] = bytes(g[a:a + 300]).lower()
        for p in rng.integers(0, len(g), size=6):
            g[int(p)] = ord("N")
        g[int(rng.integers(0, len(g)))] = ord("x")
    genomes = [(t, bytes(g)) for (t, _), g in zip(genomes, gs)]
    # two more genomes that only take part in the update (not in the fill): a chimera of genomes 0 and 3 filed under
    # genome 1's species (values move to the genus / the root), and a copy of genome 2 filed under its own species (no change)
    chim = bytes(gs[0][2000:9000]) + b"NN" + bytes(gs[3][500:6000])
    extra = [(genomes[1][0], chim), (genomes[2][0], bytes(gs[2]))]
    fill = [True] * len(genomes) + [False, False]
    return nodes, names, genomes + extra, fill


def test_db_update_matches_oracle_full_build(oracle, native, gpu_ctx):
    nodes, names, genomes, fill = _project()
    odb_full = oracle.OracleDb.build(K, nodes, names, genomes, fill=fill)
    odb_fill = oracle.OracleDb.build(K, nodes, names, genomes, fill=fill, skip_update=True)
    gdb = util.upload(oracle, native, gpu_ctx, odb_fill)
    try:
        keys_f, vals_f = odb_fill.export()
        keys_u, vals_u = odb_full.export()
        np.testing.assert_array_equal(keys_f, keys_u)
        assert (vals_f != vals_u).sum() > 100, "the update phase must have something to do in this project"
        np.testing.assert_array_equal(gdb.values(), vals_f)
        taxids = odb_fill.taxids()
        vidx_of = {t: i for i, t in enumerate(taxids)}
        seq = np.frombuffer(b"".join(g for _, g in genomes), dtype=np.uint8)
        offsets = np.zeros(len(genomes) + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([len(g) for _, g in genomes])
        vidx = np.array([vidx_of[t] for t, _ in genomes], dtype=np.int32)
        # in two calls and in a scrambled region order: the result does not depend on it (FastaReaderGoal.java:104-108)
        order = [5, 2, 6, 0]
        rest = [1, 3, 4]
        changed = 0
        for sel in (order, rest):
            s = np.concatenate([seq[int(offsets[i]):int(offsets[i + 1])] for i in sel])
            o = np.zeros(len(sel) + 1, dtype=np.uint64)
            o[1:] = np.cumsum([int(offsets[i + 1] - offsets[i]) for i in sel])
            changed += gdb.update(s, o, vidx[sel])
        assert changed >= (vals_f != vals_u).sum()
        np.testing.assert_array_equal(gdb.values(), vals_u)
        # idempotent
        assert gdb.update(seq, offsets, vidx) == 0
        np.testing.assert_array_equal(gdb.values(500, 1000), vals_u[500:1500])
        # the probe table was rebuilt with the new values: lookups and a match run agree with the oracle's full build
        v, p = gdb.lookup(keys_u[::5], use_bloom=False)
        exp = [odb_full.get(int(x)) for x in keys_u[::5]]
        np.testing.assert_array_equal(v, np.array([e[0] for e in exp], dtype=np.int32))
        bases, offs, src = synth.sample_reads([g.upper() for _, g in genomes[:5]], 3000, 150, seed=5, frac_db=0.8, sub_rate=0.01, n_rate=0.002)
        bases[bases == 0] = ord("N")  # reverse-strand copies of the genomes' non-CGAT bytes (a NUL would be dropped by the FASTQ reader)
        fq = synth.fastq_bytes(bases, offs, src)
        for cfg in (dict(), dict(layout=1)):
            orun = odb_full.match_files(util.oracle_cfg(oracle, K, **cfg), [fq])
            res, ev, counts, top, _, _ = util.gpu_match(native, gdb, bases, offs, batch=1000, **cfg)
            util.assert_match_parity(native, orun, res, counts, top)
    finally:
        gdb.close()
        odb_fill.free()
        odb_full.free()


def test_db_update_argument_checks(oracle, native, gpu_ctx):
    nodes, names, genomes, fill = _project()
    odb = oracle.OracleDb.build(K, nodes, names, genomes[:5], skip_update=True)
    gdb = util.upload(oracle, native, gpu_ctx, odb)
    try:
        seq = np.frombuffer(genomes[0][1], dtype=np.uint8)
        with pytest.raises(native.GenestripError):
            gdb.update(seq, np.array([0, 10], dtype=np.uint64), np.array([0], dtype=np.int32))     # offsets do not span the buffer
        with pytest.raises(native.GenestripError):
            gdb.update(seq, np.array([0, 50, 20, len(seq)], dtype=np.uint64), np.array([0, 0, 0], dtype=np.int32))
        # a region without node (< 0) and a region shorter than k are skipped
        before = gdb.values()
        assert gdb.update(seq, np.array([0, 20, len(seq)], dtype=np.uint64), np.array([1, -1], dtype=np.int32)) == 0
        np.testing.assert_array_equal(gdb.values(), before)
        # a session that leases the table's seen bits blocks the update
        sess = native.MatchSession(gdb, native.default_match_cfg())
        try:
            with pytest.raises(native.GenestripError):
                gdb.update(seq, np.array([0, len(seq)], dtype=np.uint64), np.array([0], dtype=np.int32))
        finally:
            sess.close()
    finally:
        gdb.close()
        odb.free()


def test_db_file_round_trip(oracle, native, gpu_ctx, tmp_path):
    """gs_db_save_file / gs_db_load_file (flat GSB1 file instead of Database.save's Java object streams, C/store/Database.java:201-314):
    an updated database written to disk and loaded again answers and matches exactly like the oracle's full build."""
    nodes, names, genomes, fill = _project()
    odb_full = oracle.OracleDb.build(K, nodes, names, genomes, fill=fill)
    odb_fill = oracle.OracleDb.build(K, nodes, names, genomes, fill=fill, skip_update=True)
    gdb = util.upload(oracle, native, gpu_ctx, odb_fill)
    path = str(tmp_path / "db.gsb")
    try:
        taxids = odb_fill.taxids()
        vidx_of = {t: i for i, t in enumerate(taxids)}
        seq = np.frombuffer(b"".join(g for _, g in genomes), dtype=np.uint8)
        offsets = np.zeros(len(genomes) + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([len(g) for _, g in genomes])
        gdb.update(seq, offsets, np.array([vidx_of[t] for t, _ in genomes], dtype=np.int32))
        gdb.save(path)
    finally:
        gdb.close()
    g2 = native.Database.load(gpu_ctx, path)
    try:
        keys_u, vals_u = odb_full.export()
        assert (g2.k, g2.n_kmers, g2.n_values) == (K, len(keys_u), odb_full.n_values)
        np.testing.assert_array_equal(g2.values(), vals_u)
        rng = np.random.default_rng(2)
        q = np.concatenate([keys_u[::3], rng.integers(0, 1 << 62, size=3000, dtype=np.int64)])
        for use_bloom in (True, False):   # the Bloom filter words travelled with the file
            v, p = g2.lookup(q, use_bloom=use_bloom)
            exp = [odb_full.get(int(x)) for x in q]
            np.testing.assert_array_equal(v, np.array([e[0] for e in exp], dtype=np.int32))
        bases, offs, src = synth.sample_reads([g.upper() for _, g in genomes[:5]], 2000, 150, seed=6, frac_db=0.8, sub_rate=0.01, n_rate=0.002)
        bases[bases == 0] = ord("N")
        fq = synth.fastq_bytes(bases, offs, src)
        for cfg in (dict(), dict(layout=1)):
            orun = odb_full.match_files(util.oracle_cfg(oracle, K, **cfg), [fq])
            res, ev, counts, top, _, _ = util.gpu_match(native, g2, bases, offs, batch=800, **cfg)
            util.assert_match_parity(native, orun, res, counts, top)
        # a truncated or foreign file is refused
        open(str(tmp_path / "bad.gsb"), "wb").write(open(path, "rb").read()[:1000])
        with pytest.raises(native.GenestripError):
            native.Database.load(gpu_ctx, str(tmp_path / "bad.gsb"))
        open(str(tmp_path / "bad2.gsb"), "wb").write(b"not a database")
        with pytest.raises(native.GenestripError):
            native.Database.load(gpu_ctx, str(tmp_path / "bad2.gsb"))
    finally:
        g2.close()
        odb_fill.free()
        odb_full.free()


def test_db_file_round_trip_keeps_values_without_a_tree_node(native, gpu_ctx, tmp_path):
    """A stored value whose tax id has no tree node matches like a miss (C/store/Database.java:136-143) but must come back from
    gs_db_get_values / a saved file with its own index, not as index 32767; and a header that lies about sizes is refused
    before anything is allocated from it."""
    rng = np.random.default_rng(11)
    n, V = 5000, 6
    keys = np.unique(rng.integers(0, 1 << 62, size=n, dtype=np.int64))
    vidx = rng.integers(0, V, size=len(keys))
    vals = (vidx - 32768).astype(np.int16)
    parent = np.array([-1, 0, 0, 1, -1, 2], dtype=np.int32)
    has_node = np.array([1, 1, 1, 1, 0, 1], dtype=np.int32)     # value index 4: taxon not in the tree
    db = native.Database(gpu_ctx, K, keys, vals, V, parent_by_vidx=parent, has_node=has_node, bloom=None, build_bloom=True)
    path = str(tmp_path / "nonode.gsb")
    try:
        np.testing.assert_array_equal(db.values(), vals)
        v, _ = db.lookup(keys)
        np.testing.assert_array_equal(v, np.where(vidx == 4, -1, vidx))
        db.save(path)
    finally:
        db.close()
    g2 = native.Database.load(gpu_ctx, path)
    try:
        np.testing.assert_array_equal(g2.values(), vals)
        v, _ = g2.lookup(keys)
        np.testing.assert_array_equal(v, np.where(vidx == 4, -1, vidx))
    finally:
        g2.close()
    raw = bytearray(open(path, "rb").read())
    for field_off, val in ((16, 1 << 40), (24, 70000), (48, 1 << 50)):   # n_kmers, n_values, bloom_words
        bad = bytearray(raw)
        bad[field_off:field_off + 8] = int(val).to_bytes(8, "little")
        open(str(tmp_path / "lie.gsb"), "wb").write(bad)
        with pytest.raises(native.GenestripError):
            native.Database.load(gpu_ctx, str(tmp_path / "lie.gsb"))
