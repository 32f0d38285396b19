"""GPU tests at the sizes BASELINE.json names, through size-independent properties (the oracle cannot run 1e8-key databases in
seconds): linearity over batches (a checksum of checksums), idempotence of the unique-k-mer bits, invariance under read
order and batch split, equality of the two device layouts, and an oracle spot check on a sample."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
K = 31


@pytest.fixture(scope="module")
def viral(native, gpu_ctx):
    import torch
    import bench
    dev = torch.device("cuda:0")
    wl = dict(bench.WORKLOADS["viral"])
    keys, vals_raw, parent, codes = bench.make_database(torch, dev, bench.DATABASES[wl["db"]], seed=43)
    db = native.Database.from_pointers(gpu_ctx, K, keys.data_ptr(), vals_raw.data_ptr(), keys.numel(), len(parent), parent, build_bloom=True)
    R = 300_000
    bases, offsets = bench.make_reads(torch, dev, wl, codes, R, seed=99)
    b_h = bases[: R * 150].cpu().numpy()
    o_h = offsets.cpu().numpy().astype(np.uint64)
    yield dict(db=db, keys=keys, vals=vals_raw, parent=parent, bases=b_h, offsets=o_h, R=R, n=int(keys.numel()))
    db.close()


def _run(native, db, bases, offsets, splits=(0, None), order=None, repeat=1, **cfg):
    sess = native.MatchSession(db, native.default_match_cfg(**cfg))
    res = []
    try:
        n = len(offsets) - 1
        bounds = [s if s is not None else n for s in splits]
        for _ in range(repeat):
            for a, b in zip(bounds[:-1], bounds[1:]):
                t = sess.submit(bases, np.ascontiguousarray(offsets[a:b + 1]), a)
                res.append(sess.collect(t)[0])
        counts, _ = sess.finish()
    finally:
        sess.close()
    return np.concatenate(res), counts


FIELDS = ("kmers", "contigs", "contig_len_squared_sum", "reads_1kmer", "reads", "reads_kmers", "reads_bps")


def test_full_size_properties(viral, native):
    db, bases, offsets, R = viral["db"], viral["bases"], viral["offsets"], viral["R"]
    res, c = _run(native, db, bases, offsets)
    assert 0.3 < c["kmers"].sum() / (R * 120.0) < 0.5                    # ~50 % of the reads come from the database
    assert (c["unique_kmers"] <= c["kmers"]).all() and c["unique_kmers"].sum() <= viral["n"]
    assert c["reads"].sum() == ((res["flags"] & native.GS_READ_ACCEPTED) != 0).sum()
    assert c["reads_kmers"].sum() == res["read_kmers"][(res["flags"] & native.GS_READ_ACCEPTED) != 0].astype(np.int64).sum()
    assert c["reads_bps"].sum() == 150 * c["reads"].sum()
    assert (c["contig_len_squared_sum"] >= c["kmers"]).all() and (c["max_contig_len"] <= 120).all()
    # batch split invariance and linearity: three batches == one batch; the same reads twice double every sum but not `unique`
    res3, c3 = _run(native, db, bases, offsets, splits=(0, 70_001, 200_000, None))
    np.testing.assert_array_equal(res3, res)
    for f in FIELDS + ("unique_kmers", "max_contig_len", "max_contig_read_no"):
        np.testing.assert_array_equal(c3[f], c[f], err_msg=f)
    _, c2 = _run(native, db, bases, offsets, repeat=2)
    for f in FIELDS:
        np.testing.assert_array_equal(c2[f], 2 * c[f], err_msg=f)
    np.testing.assert_array_equal(c2["unique_kmers"], c["unique_kmers"])
    np.testing.assert_array_equal(c2["max_contig_len"], c["max_contig_len"])
    # the reference's own structures on the device give the same answers as the probe table
    res_c, c_c = _run(native, db, bases, offsets, layout=native.GS_LAYOUT_CLASSIC)
    np.testing.assert_array_equal(res_c, res)
    for f in FIELDS + ("unique_kmers", "max_contig_len", "max_contig_read_no"):
        np.testing.assert_array_equal(c_c[f], c[f], err_msg=f)
    _, c_nb = _run(native, db, bases, offsets, layout=native.GS_LAYOUT_CLASSIC, use_bloom_filter=0)
    for f in FIELDS + ("unique_kmers",):
        np.testing.assert_array_equal(c_nb[f], c[f], err_msg=f)


def test_read_order_invariance_of_taxon_sums(viral, native):
    db, bases, offsets, R = viral["db"], viral["bases"], viral["offsets"], viral["R"]
    n = 100_000
    _, c = _run(native, db, bases, offsets[: n + 1])
    perm = np.random.default_rng(1).permutation(n)
    pb = bases[: n * 150].reshape(n, 150)[perm].reshape(-1).copy()
    res_p, c_p = _run(native, db, pb, offsets[: n + 1])
    for f in FIELDS + ("unique_kmers", "max_contig_len"):
        np.testing.assert_array_equal(c_p[f], c[f], err_msg=f)


def test_lookup_round_trip_at_full_size(viral, native):
    import torch
    db, keys = viral["db"], viral["keys"]
    idx = torch.randint(0, keys.numel(), (200_000,), device=keys.device)
    q = keys[idx].cpu().numpy()
    v, p = db.lookup(q)                      # checks the probe table against the sorted array on every query as well
    np.testing.assert_array_equal(p, idx.cpu().numpy())
    np.testing.assert_array_equal(v, (viral["vals"][idx].to(torch.int32) + 32768).cpu().numpy())
    miss = np.random.default_rng(2).integers(0, 1 << 62, size=200_000, dtype=np.int64)
    v, p = db.lookup(miss)
    present = np.isin(miss, q)
    assert (v[~present] == -1).mean() > 0.9999


def test_oracle_spot_check_at_full_size(viral, native, oracle):
    """The CPU oracle on the full 1e8-key database, 20 000 reads: every per-read and per-taxon integer must agree."""
    import util
    n = 20_000
    keys_h = viral["keys"].cpu().numpy()
    vals_h = viral["vals"].cpu().numpy()
    odb = oracle.OracleDb.from_arrays(K, keys_h, vals_h, len(viral["parent"]), viral["parent"], build_bloom=True)
    try:
        off = viral["offsets"][: n + 1]
        bb = viral["bases"][: n * 150].tobytes()
        fq = b"".join(b"@r%d x\n%s\n+\n%s\n" % (i, bb[i * 150:(i + 1) * 150], b"I" * 150) for i in range(n))
        orun = odb.match_files(oracle.match_cfg(k=K), [fq])
        res, c = _run(native, viral["db"], viral["bases"], off)
        util.assert_match_parity(native, orun, res, c)
    finally:
        odb.free()
