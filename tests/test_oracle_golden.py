"""The CPU oracle against the reference's own known-answer tests, fixtures and golden vectors.

The reference is Java and cannot run here (no JDK); these are its unit tests restated statement by statement in
oracle/gs_oracle_kat.cpp, with the reference test file:line cited there.  This is what pins the oracle.
"""
import gzip
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _kat(oracle, name, *args):
    L = oracle.lib()
    fails = getattr(L, name)(*args)
    assert fails == 0, L.gso_kat_message().decode()


def test_java_random_check_values(oracle):
    _kat(oracle, "gso_kat_random")
    out = np.zeros(3, dtype=np.int64)
    oracle.lib().gso_random_longs(42, 3, out.ctypes.data)
    assert list(out) == [-5025562857975149833, -5843495416241995736, 5694868678511409995]


def test_next_kmer_rolling_equals_full_encode(oracle):  # T/util/NextKMerTest.java:37-87
    _kat(oracle, "gso_kat_next_kmer", 1000)


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["blocked", "xor", "murmur"])
def test_bloom_no_false_negatives_and_fpp(oracle, kind):  # T/bloom/KMerBloomFilterTest.java:47-130 (size 5*100*1000)
    _kat(oracle, "gso_kat_bloom", kind, 500000)


def test_xor_bloom_long_min_value_and_sizing(oracle):  # T/bloom/XORKMerBloomFilterTest.java:50-58
    _kat(oracle, "gso_kat_xor_min_value")


@pytest.mark.parametrize("store", [0, 2], ids=["sorted", "radix"])
def test_store_put_get_visit(oracle, store):  # T/store/AbstractKMerStoreTest.java:118-261 (reduced from 1 M to 200 k entries)
    _kat(oracle, "gso_kat_store", store, 200000, 200000)


def test_sorted_and_radix_stores_agree(oracle):  # T/match/RadixKMerStoreBenchmarkTest.java:186-229
    _kat(oracle, "gso_kat_cross_store", 50000)


@pytest.mark.parametrize("store,optimize", [(0, 0), (0, 1), (2, 0), (2, 1)])
def test_match_read_kat(oracle, store, optimize):  # T/match/FastqKMerMatcherTest.java:96-210
    _kat(oracle, "gso_kat_match_read", store, optimize)


@pytest.mark.parametrize("store", [0, 2], ids=["sorted", "radix"])
def test_read_classification_table(oracle, store):  # T/match/FastqKMerMatcherTest.java:315-412
    _kat(oracle, "gso_kat_classification", store)


def test_small_tax_tree_lca(oracle):  # T/tax/SmallTaxTreeLCATest.java:56-125
    _kat(oracle, "gso_kat_lca")


@pytest.mark.parametrize("with_probs", [0, 1])
def test_fastq_reader_fixture(oracle, with_probs):  # T/fastq/FastqReaderTest.java:38-75, R/fastq/SimpleTest.fastq
    _kat(oracle, "gso_kat_fastq_reader", with_probs)


def test_simple_fastq_fixture_file_matches_embedded_copy(oracle):
    """tests/golden/SimpleTest.fastq is the reference fixture R/fastq/SimpleTest.fastq (a data file, 19 lines)."""
    path = os.path.join(HERE, "golden", "SimpleTest.fastq")
    data = open(path, "rb").read()
    flt = oracle.Bloom(kind=0)
    flt.ensure(10)
    run = oracle.filter_files(flt, 2, [data], with_probs=True, initial_read_size=3)
    flt.free()
    assert run.n_reads == 2
    assert list(run.read_size) == [60, 4]
    # rewritten records carry the joined read and quality strings the reference test expects
    lines = run.rest.split(b"\n")
    assert lines[1] == b"GATTTGGGGTTCAAAGCAGTATCGATCAAATAGTAAATCCATTTGTTCAACTCACAGTTT"
    assert lines[3] == b"!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65"
    assert lines[5] == b"CGAT" and lines[7] == b"!**>"


def test_sample_fastq_totals(oracle):
    """ref/README.md:169 TOTAL row of the sample run: 6565 reads, 658255 bp, 461305 31-mers.  The golden file holds the
    read lengths of ref/data/projects/human_virus/fastq/sample.fastq.gz (made by tests/golden/make_golden.py)."""
    lens = np.loadtxt(os.path.join(HERE, "golden", "sample_fastq_read_lengths.txt.gz"), dtype=np.int64)
    assert len(lens) == 6565 and lens.sum() == 658255 and np.maximum(lens - 30, 0).sum() == 461305
    # the oracle's parser reproduces these totals on a FASTQ rebuilt with the same lengths
    fq = b"".join(b"@s%d\n%s\n+\n%s\n" % (i, b"A" * int(n), b"I" * int(n)) for i, n in enumerate(lens))
    flt = oracle.Bloom(kind=0)
    flt.ensure(10)
    run = oracle.filter_files(flt, 31, [fq])
    flt.free()
    assert (run.total_reads, run.total_bps, run.total_kmers) == (6565, 658255, 461305)


def test_sample_fastq_itself_through_the_oracle_parser(oracle):
    """The same totals from the reference's sample file itself (real headers, qualities, N runs): the oracle's restatement of
    AbstractFastqReader.doReadFastq + BufferedLineReader on real data, and the product's host parser on the same bytes."""
    import gzip
    text = gzip.decompress(open(os.path.join(HERE, "golden", "human_virus_sample.fastq.gz"), "rb").read())
    flt = oracle.Bloom(kind=0)
    flt.ensure(10)
    run = oracle.filter_files(flt, 31, [text], with_probs=True)
    flt.free()
    assert (run.total_reads, run.total_bps, run.total_kmers) == (6565, 658255, 461305)
    assert run.rest == text   # nothing accepted: every record is rewritten, byte for byte, to the rest stream


def test_kraken_line_shape(oracle):
    """R/projects/dengue1/test.out: `C\\ttest\\t1\\t41\\t0:2 1:7 0:2` -- shape of writeMatchDetails + printKrakenStyleOut."""
    golden = open(os.path.join(HERE, "golden", "dengue1_test.out"), "rb").read()
    assert golden == b"C\ttest\t1\t41\t0:2 1:7 0:2\n"
    # 41 bases, k=31 -> 11 k-mers: 2 misses, 7 hits of taxon 1, 2 misses
    rng = np.random.default_rng(1)
    from genestrip_b200 import synth
    g = synth.random_genome(rng, 37).tobytes()           # 7 consecutive 31-mers
    nodes = "1\t|\t1\t|\tno rank\t|\t\t|\n"
    names = "1\t|\troot\t|\t\t|\tscientific name\t|\n"
    odb = oracle.OracleDb.build(31, nodes, names, [("1", g)])
    read = b"TG" + g + b"CA"
    run = odb.match_files(oracle.match_cfg(k=31, write_kraken=True), [b"@test\n" + read + b"\n+\n" + b"I" * 41 + b"\n"])
    odb.free()
    assert run.kraken == golden


def test_java_double_to_string(oracle):
    cases = {1.0: "1.0", 0.5: "0.5", 100.0: "100.0", 1234567.0: "1234567.0", 1.0e7: "1.0E7", 1.0e-3: "0.001", 1.0e-4: "1.0E-4",
             0.1: "0.1", 1 / 3: "0.3333333333333333", 123456789.125: "1.23456789125E8", 2.0e-5: "2.0E-5", 150.0: "150.0"}
    for v, s in cases.items():
        assert oracle.java_double_to_string(v) == s


def dengue1_project():
    """T/goals/refseq/DBGoalTest.java:76-142 without the network: the DENV-1 genome (R/projects/dengue1/dengue1.fasta,
    = RefSeq NC_001477.1) filled under tax id 11053 (taxids.txt), and the same file registered again under 9606
    (additional.txt, "obviously wrong and just for the update test") so that the update phase moves every k-mer to
    LCA(11053, 9606) = 1.  Lineage ranks as in NCBI; only the tree shape matters."""
    lineage = [(1, 1, "no rank"), (10239, 1, "superkingdom"), (11050, 10239, "family"), (11051, 11050, "genus"),
               (12637, 11051, "species"), (11053, 12637, "no rank"), (131567, 1, "no rank"), (9606, 131567, "species")]
    nodes = "".join("%d\t|\t%d\t|\t%s\t|\t\t|\n" % e for e in lineage)
    names = "".join("%d\t|\tn%d\t|\t\t|\tscientific name\t|\n" % (e[0], e[0]) for e in lineage)
    fasta = open(os.path.join(HERE, "golden", "dengue1.fasta"), "rb").read()
    fastq = open(os.path.join(HERE, "golden", "dengue1_test.fastq"), "rb").read()
    golden = open(os.path.join(HERE, "golden", "dengue1_test.out"), "rb").read()
    return nodes, names, fasta, fastq, golden


def test_dengue1_golden_kraken_output(oracle):
    """The reference's own end-to-end golden vector for this path: R/projects/dengue1/test.out must be reproduced
    byte for byte (DBGoalTest.testKrakenOutput, T/goals/refseq/DBGoalTest.java:127-142)."""
    nodes, names, fasta, fastq, golden = dengue1_project()
    odb = oracle.OracleDb.build(31, nodes, names, [("11053", fasta), ("9606", fasta)], requested=["11053"], fill=[True, False])
    try:
        taxids = odb.taxids()
        dbk = odb.db_kmers()
        # DBGoalTest.java:102-111: after the update no k-mer is left on 11053, all are on "1"
        assert dbk[taxids.index("11053")] == 0 and dbk[taxids.index("1")] == odb.n_kmers
        run = odb.match_files(oracle.match_cfg(k=31, write_kraken=True), [fastq])
        assert run.kraken == golden
    finally:
        odb.free()
