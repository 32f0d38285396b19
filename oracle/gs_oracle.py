"""ctypes wrapper of the CPU ORACLE (oracle/gs_oracle*.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package genestrip_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libgs_oracle.so")
_P = C.c_void_p


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("gs_oracle_capi.cpp", "gs_oracle_kat.cpp", "gs_oracle.hpp")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    res = subprocess.run(["make", "-C", HERE, "-B"] if force else ["make", "-C", HERE], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


class MatchCfg(C.Structure):
    _fields_ = [("k", C.c_int), ("classify", C.c_int), ("max_paths", C.c_int),
                ("max_read_tax_err", C.c_double), ("max_read_class_err", C.c_double),
                ("threshold", C.c_int), ("max_kmer_res_counts", C.c_int), ("count_unique", C.c_int), ("write_all", C.c_int),
                ("write_kraken", C.c_int), ("write_filtered", C.c_int), ("with_probs", C.c_int),
                ("initial_read_size", C.c_int), ("use_filter", C.c_int), ("dump_labels", C.c_int)]


def match_cfg(k=31, classify=True, max_paths=10, max_read_tax_err=-1.0, max_read_class_err=-1.0, threshold=1,
              max_kmer_res_counts=0, count_unique=True, write_all=True, write_kraken=False, write_filtered=False,
              with_probs=False, initial_read_size=4096, use_filter=True, dump_labels=False):
    return MatchCfg(k, int(classify), max_paths, max_read_tax_err, max_read_class_err, threshold, max_kmer_res_counts,
                    int(count_unique), int(write_all), int(write_kraken), int(write_filtered), int(with_probs),
                    initial_read_size, int(use_filter), int(dump_labels))


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.gso_kat_message.restype = C.c_char_p
        L.gso_db_new.restype = _P
        L.gso_db_new.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_double]
        L.gso_db_free.argtypes = [_P]
        L.gso_db_error.restype = C.c_char_p
        L.gso_db_error.argtypes = [_P]
        L.gso_db_request.argtypes = [_P, C.c_char_p]
        L.gso_db_add_genome.argtypes = [_P, C.c_char_p, _P, C.c_size_t, C.c_int]
        L.gso_db_finalize.argtypes = [_P, C.c_double, C.c_int]
        L.gso_db_from_arrays.restype = _P
        L.gso_db_from_arrays.argtypes = [C.c_int, _P, _P, C.c_int64, C.c_int, _P, C.c_int]
        for f in ("gso_db_k", "gso_db_n_values"):
            getattr(L, f).argtypes = [_P]
        L.gso_db_n_kmers.restype = C.c_int64
        L.gso_db_n_kmers.argtypes = [_P]
        L.gso_db_value_taxid.restype = C.c_char_p
        L.gso_db_value_taxid.argtypes = [_P, C.c_int]
        L.gso_db_export.argtypes = [_P, _P, _P]
        L.gso_db_tree.argtypes = [_P, _P, _P, _P, _P]
        L.gso_db_dbkmers.argtypes = [_P, _P]
        L.gso_db_store_filter.restype = _P
        L.gso_db_store_filter.argtypes = [_P]
        L.gso_db_index_filter.restype = _P
        L.gso_db_index_filter.argtypes = [_P]
        L.gso_db_set_use_filter.argtypes = [_P, C.c_int]
        L.gso_db_get.argtypes = [_P, C.c_int64, C.POINTER(C.c_int64)]
        L.gso_db_node_name.restype = C.c_char_p
        L.gso_db_node_name.argtypes = [_P, C.c_int]
        L.gso_db_node_rank.argtypes = [_P, C.c_int]
        L.gso_db_node_requested.argtypes = [_P, C.c_int]
        L.gso_bloom_new.restype = _P
        L.gso_bloom_new.argtypes = [C.c_int, C.c_double]
        L.gso_bloom_free.argtypes = [_P]
        L.gso_bloom_ensure.restype = C.c_int64
        L.gso_bloom_ensure.argtypes = [_P, C.c_int64]
        L.gso_bloom_put.argtypes = [_P, _P, C.c_int64]
        L.gso_bloom_contains.argtypes = [_P, _P, C.c_int64, _P]
        L.gso_bloom_kind.argtypes = [_P]
        L.gso_bloom_params.argtypes = [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.gso_bloom_words.restype = _P
        L.gso_bloom_words.argtypes = [_P]
        L.gso_bloom_factors.restype = _P
        L.gso_bloom_factors.argtypes = [_P]
        L.gso_bloom_from_words.restype = _P
        L.gso_bloom_from_words.argtypes = [C.c_int, C.c_int64, C.c_int, _P, _P, C.c_int64]
        L.gso_filter_reads_mt.restype = C.c_int64
        L.gso_filter_reads_mt.argtypes = [_P, C.c_int, C.c_int, C.c_double, _P, _P, C.c_int64, C.c_int, _P]
        L.gso_match_files.restype = _P
        L.gso_match_files.argtypes = [_P, C.POINTER(MatchCfg), _P, _P, _P, C.c_int]
        L.gso_filter_files.restype = _P
        L.gso_filter_files.argtypes = [_P, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _P, _P, _P, C.c_int]
        L.gso_run_free.argtypes = [_P]
        L.gso_run_text.restype = _P
        L.gso_run_text.argtypes = [_P, C.c_int, C.POINTER(C.c_size_t)]
        L.gso_run_totals.argtypes = [_P, _P]
        L.gso_run_n_reads.restype = C.c_int64
        L.gso_run_n_reads.argtypes = [_P]
        L.gso_run_stat.restype = _P
        L.gso_run_stat.argtypes = [_P, C.c_int]
        L.gso_run_has_stats.restype = _P
        L.gso_run_has_stats.argtypes = [_P]
        L.gso_run_dstat.restype = _P
        L.gso_run_dstat.argtypes = [_P, C.c_int]
        L.gso_run_desc.restype = C.c_char_p
        L.gso_run_desc.argtypes = [_P, C.c_int]
        L.gso_run_read.restype = _P
        L.gso_run_read.argtypes = [_P, C.c_int]
        L.gso_run_accept.restype = _P
        L.gso_run_accept.argtypes = [_P]
        L.gso_run_n_labels.restype = C.c_int64
        L.gso_run_n_labels.argtypes = [_P]
        L.gso_run_labels.restype = _P
        L.gso_run_labels.argtypes = [_P]
        L.gso_run_label_pos.restype = _P
        L.gso_run_label_pos.argtypes = [_P]
        L.gso_run_max_counts.argtypes = [_P, C.POINTER(_P), C.POINTER(_P)]
        L.gso_match_reads_mt.restype = C.c_int64
        L.gso_match_reads_mt.argtypes = [_P, C.POINTER(MatchCfg), _P, _P, C.c_int64, C.c_int, _P]
        L.gso_all_kmers.argtypes = [_P, C.c_int, C.c_int, _P]
        L.gso_kmer_canonical.restype = C.c_int64
        L.gso_kmer_canonical.argtypes = [_P, C.c_int, C.c_int]
        L.gso_random_longs.argtypes = [C.c_int64, C.c_int, _P]
        L.gso_murmur64.restype = C.c_int64
        L.gso_murmur64.argtypes = [C.c_int64, C.c_int64]
        L.gso_java_double_to_string.argtypes = [C.c_double, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def _np_from(ptr, dtype, n):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


class Bloom:
    """A KMerProbFilter of the oracle (kind 0 blocked, 1 xor, 2 murmur); owned=False for filters inside a db."""

    def __init__(self, handle=None, kind=0, fpp=0.01, owned=True):
        self.h = handle if handle is not None else lib().gso_bloom_new(kind, fpp)
        self.owned = owned and handle is None

    @classmethod
    def from_words(cls, kind, bits, hashes, factors, words):
        """A hashed (XOR = 1 / Murmur = 2) filter from its serialized state: bit count, hash factors, bit vector words."""
        factors = np.ascontiguousarray(factors, dtype=np.int64)
        words = np.ascontiguousarray(words, dtype=np.int64)
        f = cls(handle=lib().gso_bloom_from_words(kind, bits, hashes, _ptr(factors), _ptr(words), len(words)))
        f.owned = True
        return f

    def accept_reads_mt(self, k, bases, offsets, threads, min_pos_count=1, pos_ratio=0.2):
        """FastqBloomFilter.isAcceptRead over pre-parsed reads on `threads` threads; returns (k-mers of the reads, accept uint8[n])."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        acc = np.zeros(len(offsets) - 1, dtype=np.uint8)
        n = lib().gso_filter_reads_mt(self.h, k, min_pos_count, pos_ratio, _ptr(bases), _ptr(offsets), len(offsets) - 1, threads, _ptr(acc))
        return n, acc

    def ensure(self, n):
        return lib().gso_bloom_ensure(self.h, n)

    def put(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        lib().gso_bloom_put(self.h, _ptr(keys), len(keys))

    def contains(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        out = np.zeros(len(keys), dtype=np.uint8)
        lib().gso_bloom_contains(self.h, _ptr(keys), len(keys), _ptr(out))
        return out

    @property
    def kind(self):
        return lib().gso_bloom_kind(self.h)

    def params(self):
        """(kind, p0, p1, factors, words): blocked p0=seed p1=buckets; hashed p0=bits p1=hashes."""
        p0, p1, nw = C.c_int64(), C.c_int64(), C.c_int64()
        lib().gso_bloom_params(self.h, C.byref(p0), C.byref(p1), C.byref(nw))
        words = _np_from(lib().gso_bloom_words(self.h), np.int64, nw.value)
        factors = None if self.kind == 0 else _np_from(lib().gso_bloom_factors(self.h), np.int64, p1.value)
        return self.kind, p0.value, p1.value, factors, words

    def free(self):
        if self.owned and self.h:
            lib().gso_bloom_free(self.h)
            self.h = None


class Run:
    """Results of one oracle match / filter run."""

    STAT_NAMES = ("kmers", "contigs", "sqsum", "maxlen", "reads1", "reads", "reads_kmers", "reads_bps", "unique")

    def __init__(self, h, n_values):
        L = lib()
        self.n_values = n_values

        def text(which):
            n = C.c_size_t()
            p = L.gso_run_text(h, which, C.byref(n))
            return C.string_at(p, n.value) if n.value else b""

        self.csv, self.filtered, self.kraken, self.rest = text(0), text(1), text(2), text(3)
        tot = np.zeros(3, dtype=np.int64)
        L.gso_run_totals(h, _ptr(tot))
        self.total_reads, self.total_kmers, self.total_bps = (int(x) for x in tot)
        n_reads = L.gso_run_n_reads(h)
        self.n_reads = n_reads
        if n_values:
            self.stats = {nm: _np_from(L.gso_run_stat(h, i), np.int64, n_values) for i, nm in enumerate(self.STAT_NAMES)}
            self.has_stats = _np_from(L.gso_run_has_stats(h), np.int32, n_values)
            self.dstats = [_np_from(L.gso_run_dstat(h, i), np.float64, n_values) for i in range(4)]
            self.desc = [L.gso_run_desc(h, v) for v in range(n_values)]
            names = ("class_vidx", "read_kmers", "tax_err", "accepted", "read_size")
            self.reads = {nm: _np_from(L.gso_run_read(h, i), np.int32, n_reads) for i, nm in enumerate(names)}
            nl = L.gso_run_n_labels(h)
            self.labels = _np_from(L.gso_run_labels(h), np.int32, nl)
            self.label_pos = _np_from(L.gso_run_label_pos(h), np.int64, nl)
            cp, hp = _P(), _P()
            n = L.gso_run_max_counts(h, C.byref(cp), C.byref(hp))
            self.max_counts_n = n
            if n:
                self.max_counts = _np_from(cp.value, np.int16, (n_values + 1) * n).reshape(n_values + 1, n)
                self.max_counts_has = _np_from(hp.value, np.int32, n_values + 1)
        else:
            self.accept = _np_from(L.gso_run_accept(h), np.uint8, n_reads)
            self.read_size = _np_from(L.gso_run_read(h, 4), np.int32, n_reads)
        L.gso_run_free(h)


def _files_args(files, is_fasta):
    n = len(files)
    bufs = [np.frombuffer(f, dtype=np.uint8) if len(f) else np.zeros(0, dtype=np.uint8) for f in files]
    ptrs = (C.c_void_p * n)(*[b.ctypes.data if len(b) else None for b in bufs])
    lens = (C.c_size_t * n)(*[len(f) for f in files])
    fa = (C.c_int * n)(*[int(x) for x in (is_fasta or [False] * n)])
    return bufs, ptrs, lens, fa, n


class OracleDb:
    """The oracle's Database: restated `db` goal over synthetic genomes, or straight from arrays."""

    def __init__(self, h):
        self.h = h
        L = lib()
        self.k = L.gso_db_k(h)
        self.n_kmers = L.gso_db_n_kmers(h)
        self.n_values = L.gso_db_n_values(h)

    @classmethod
    def build(cls, k, nodes_dmp, names_dmp, genomes, requested=(), use_radix=False, radix_bits=17, opt_fpp=0.01,
              index_fpp=0.0, index_xor=True, fill=None, skip_update=False):
        """genomes: list of (taxid str, bytes sequence).  fill[i]=False -> genome only takes part in the LCA update."""
        L = lib()
        nodes_dmp = nodes_dmp if isinstance(nodes_dmp, bytes) else nodes_dmp.encode()
        names_dmp = names_dmp if isinstance(names_dmp, bytes) else names_dmp.encode()
        h = L.gso_db_new(k, nodes_dmp, len(nodes_dmp), names_dmp, len(names_dmp), int(use_radix), radix_bits, opt_fpp)
        for t in requested:
            if L.gso_db_request(h, str(t).encode()) != 0:
                raise ValueError("unknown requested taxid %s" % t)
        for i, (taxid, seq) in enumerate(genomes):
            b = np.frombuffer(seq, dtype=np.uint8)
            if L.gso_db_add_genome(h, str(taxid).encode(), _ptr(b), len(b), 1 if (fill is None or fill[i]) else 0) != 0:
                raise ValueError("unknown genome taxid %s" % taxid)
        if skip_update:  # database as the fill phase leaves it (the GPU update phase is then checked against a full build)
            L.gso_db_set_skip_update.argtypes = [_P, C.c_int]
            L.gso_db_set_skip_update(h, 1)
        if L.gso_db_finalize(h, index_fpp, int(index_xor)) != 0:
            raise RuntimeError(L.gso_db_error(h).decode())
        return cls(h)

    @classmethod
    def from_arrays(cls, k, keys, vals_raw, n_values, parent_by_vidx, build_bloom=True):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        vals_raw = np.ascontiguousarray(vals_raw, dtype=np.int16)
        parent = np.ascontiguousarray(parent_by_vidx, dtype=np.int32)
        return cls(lib().gso_db_from_arrays(k, _ptr(keys), _ptr(vals_raw), len(keys), n_values, _ptr(parent), int(build_bloom)))

    def export(self):
        """(keys int64[n], vals_raw int16[n]) in storage-position order."""
        keys = np.zeros(self.n_kmers, dtype=np.int64)
        vals = np.zeros(self.n_kmers, dtype=np.int16)
        lib().gso_db_export(self.h, _ptr(keys), _ptr(vals))
        return keys, vals

    def tree(self):
        """(parent, depth, position, has_node) by value index."""
        V = self.n_values
        a = [np.zeros(V, dtype=np.int32) for _ in range(4)]
        lib().gso_db_tree(self.h, *[_ptr(x) for x in a])
        return tuple(a)

    def db_kmers(self):
        out = np.zeros(self.n_values, dtype=np.int64)
        lib().gso_db_dbkmers(self.h, _ptr(out))
        return out

    def taxids(self):
        return [lib().gso_db_value_taxid(self.h, v).decode() for v in range(self.n_values)]

    def node_names(self):
        return [lib().gso_db_node_name(self.h, v).decode() for v in range(self.n_values)]

    def node_ranks(self):
        return [lib().gso_db_node_rank(self.h, v) for v in range(self.n_values)]

    def store_filter(self):
        h = lib().gso_db_store_filter(self.h)
        return Bloom(handle=h, owned=False) if h else None

    def index_filter(self):
        h = lib().gso_db_index_filter(self.h)
        return Bloom(handle=h, owned=False) if h else None

    def get(self, kmer):
        pos = C.c_int64(-1)
        v = lib().gso_db_get(self.h, int(kmer), C.byref(pos))
        return v, pos.value

    def match_files(self, cfg, files, is_fasta=None):
        bufs, ptrs, lens, fa, n = _files_args(files, is_fasta)
        h = lib().gso_match_files(self.h, C.byref(cfg), ptrs, lens, fa, n)
        return Run(h, self.n_values)

    def match_reads_mt(self, cfg, bases, offsets, threads):
        """Multi-threaded matchRead over pre-parsed reads (CPU baseline); returns (kmers processed, kmers per value index)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        per = np.zeros(self.n_values, dtype=np.int64)
        n = lib().gso_match_reads_mt(self.h, C.byref(cfg), _ptr(bases), _ptr(offsets), len(offsets) - 1, threads, _ptr(per))
        return n, per

    def free(self):
        if self.h:
            lib().gso_db_free(self.h)
            self.h = None


def filter_files(bloom, k, files, min_pos_count=1, pos_ratio=0.2, with_probs=False, initial_read_size=4096, is_fasta=None):
    bufs, ptrs, lens, fa, n = _files_args(files, is_fasta)
    h = lib().gso_filter_files(bloom.h, k, min_pos_count, pos_ratio, int(with_probs), initial_read_size, ptrs, lens, fa, n)
    return Run(h, 0)


def all_kmers(seq, k):
    """Canonical k-mer of every window of seq (invalid windows -> -1)."""
    b = np.frombuffer(seq, dtype=np.uint8)
    out = np.zeros(max(len(b) - k + 1, 0), dtype=np.int64)
    if len(out):
        lib().gso_all_kmers(_ptr(b), len(b), k, _ptr(out))
    return out


def java_double_to_string(v):
    buf = C.create_string_buffer(64)
    n = lib().gso_java_double_to_string(v, buf, 64)
    return buf.value[:n].decode()
