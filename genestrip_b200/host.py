"""ctypes binding of the C++ host layer (csrc/gs_host.*): the `match` and `filter` goals as whole-file operations.

This is what a user of the reference calls instead of GSMaker.match()/filter(): FASTQ/FASTA in, CSV + filtered FASTQ +
kraken-style lines out; parsing, batching, result ordering and CSV writing are host C++, the per-read work is CUDA.
"""
import ctypes as C

import numpy as np

from . import capi

_P = C.c_void_p
_ready = False


class HostMatchCfg(C.Structure):
    _fields_ = [("classify_reads", C.c_int), ("count_unique_kmers", C.c_int), ("use_bloom_filter", C.c_int),
                ("max_kmer_res_counts", C.c_int), ("max_classification_paths", C.c_int), ("min_kmers_for_class", C.c_int),
                ("max_read_tax_error_count", C.c_double), ("max_read_class_error_count", C.c_double),
                ("write_all", C.c_int), ("with_probs", C.c_int), ("initial_read_size_bytes", C.c_int), ("layout", C.c_int),
                ("write_filtered", C.c_int), ("write_kraken", C.c_int), ("batch_reads", C.c_uint32), ("text_chunk_bytes", C.c_uint32)]


def _lib():
    global _ready
    L = capi.lib()
    if not _ready:
        L.gsh_meta_new.restype = _P
        L.gsh_meta_new.argtypes = [C.c_int, C.c_int, C.c_int64]
        L.gsh_meta_free.argtypes = [_P]
        L.gsh_meta_set_node.argtypes = [_P, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]
        L.gsh_match_goal.restype = _P
        L.gsh_match_goal.argtypes = [_P, _P, C.POINTER(HostMatchCfg), _P, _P, _P, _P, C.c_int, C.c_char_p, C.c_char_p]
        L.gsh_filter_goal.restype = _P
        L.gsh_filter_goal.argtypes = [_P, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint32, _P, _P, _P, _P, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_uint32]
        L.gsh_parse_only.restype = _P
        L.gsh_parse_only.argtypes = [C.c_int, C.c_int, _P, _P, _P, _P, C.c_int]
        L.gsh_result_free.argtypes = [_P]
        L.gsh_result_error.restype = C.c_char_p
        L.gsh_result_error.argtypes = [_P]
        L.gsh_result_text.restype = _P
        L.gsh_result_text.argtypes = [_P, C.c_int, C.POINTER(C.c_size_t)]
        L.gsh_result_totals.argtypes = [_P, _P]
        L.gsh_result_accept.restype = _P
        L.gsh_result_accept.argtypes = [_P, C.POINTER(C.c_size_t)]
        L.gsh_result_dsums.restype = _P
        L.gsh_result_dsums.argtypes = [_P, C.POINTER(C.c_size_t)]
        L.gsh_result_launches.restype = C.c_uint64
        L.gsh_result_launches.argtypes = [_P]
        L.gsh_result_feeder.argtypes = [_P, _P]
        L.gsh_last_record_start.restype = C.c_size_t
        L.gsh_last_record_start.argtypes = [_P, C.c_size_t]
        L.gsh_java_double_to_string.argtypes = [C.c_double, C.c_char_p, C.c_int]
        L.gsh_device_inflated_blocks.restype = C.c_uint64
        L.gsh_device_inflated_blocks.argtypes = []
        L.gsh_bgzf_read_all.restype = _P
        L.gsh_bgzf_read_all.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int64)]
        _ready = True
    return L


class DbMeta:
    """Names, ranks, tree positions and per-taxon database k-mer counts by value index (Database / SmallTaxTree metadata)."""

    def __init__(self, k, total_kmers, taxids, names, ranks, parent, position, level, has_node, db_kmers):
        L = _lib()
        n = len(taxids)
        self.n_values = n
        self.h = L.gsh_meta_new(k, n, int(total_kmers))
        for v in range(n):
            L.gsh_meta_set_node(self.h, v, str(taxids[v]).encode(), (names[v] or "").encode(), int(ranks[v]), int(parent[v]),
                                int(position[v]), int(level[v]), int(has_node[v]), int(db_kmers[v]))

    def free(self):
        if self.h:
            _lib().gsh_meta_free(self.h)
            self.h = None


class GoalResult:
    def __init__(self, h):
        L = _lib()
        err = L.gsh_result_error(h).decode()
        try:
            if err:
                raise capi.GenestripError(-1, err)

            def text(which):
                n = C.c_size_t()
                p = L.gsh_result_text(h, which, C.byref(n))
                return C.string_at(p, n.value) if n.value else b""

            self.csv, self.filtered, self.kraken, self.rest = text(0), text(1), text(2), text(3)
            tot = np.zeros(3, dtype=np.int64)
            L.gsh_result_totals(h, tot.ctypes.data_as(_P))
            self.total_reads, self.total_kmers, self.total_bps = (int(x) for x in tot)
            n = C.c_size_t()
            p = L.gsh_result_accept(h, C.byref(n))
            self.accept = np.frombuffer(C.string_at(p, n.value), dtype=np.uint8).copy() if n.value else np.zeros(0, np.uint8)
            p = L.gsh_result_dsums(h, C.byref(n))
            self.dsums = np.frombuffer(C.string_at(p, n.value * 8), dtype=np.float64).copy().reshape(4, -1) if n.value else None
            self.launches = L.gsh_result_launches(h)
            fd = np.zeros(2, dtype=np.uint64)
            L.gsh_result_feeder(h, fd.ctypes.data_as(_P))
            self.text_chunks, self.text_chunks_refused = int(fd[0]), int(fd[1])
        finally:
            L.gsh_result_free(h)


def _inputs(files, is_fasta):
    n = len(files)
    keep = []
    data = (C.c_void_p * n)()
    lens = (C.c_size_t * n)()
    paths = (C.c_char_p * n)()
    fa = (C.c_int * n)(*[int(x) for x in (is_fasta or [False] * n)])
    for i, f in enumerate(files):
        if isinstance(f, str):
            paths[i] = f.encode()
        else:
            b = np.frombuffer(f, dtype=np.uint8) if len(f) else np.zeros(1, dtype=np.uint8)
            keep.append(b)
            data[i] = b.ctypes.data
            lens[i] = len(f)
            paths[i] = None
    return keep, data, lens, paths, fa, n


def match_goal(db, meta, files, is_fasta=None, write_filtered=False, write_kraken=False, filtered_path=None, kraken_path=None,
               batch_reads=0, gpu_parse=True, text_chunk_bytes=0, **cfg):
    """`match` for one key over files (bytes or paths).  cfg: gs_match_cfg style keywords + write_all / with_probs / initial_read_size_bytes.
    gpu_parse: FASTQ inputs go to the GPU as raw text chunks of text_chunk_bytes (the device splits the records); chunks that
    are not strict 4-line FASTQ fall back to the sequential host parser."""
    c = HostMatchCfg(cfg.get("classify_reads", 1), cfg.get("count_unique_kmers", 1), cfg.get("use_bloom_filter", 1),
                     cfg.get("max_kmer_res_counts", 0), cfg.get("max_classification_paths", 10), cfg.get("min_kmers_for_class", 1),
                     cfg.get("max_read_tax_error_count", -1.0), cfg.get("max_read_class_error_count", -1.0),
                     int(cfg.get("write_all", 1)), int(cfg.get("with_probs", 0)), cfg.get("initial_read_size_bytes", 4096),
                     cfg.get("layout", 0), int(write_filtered), int(write_kraken), batch_reads,
                     (int(text_chunk_bytes) if gpu_parse else 0xFFFFFFFF))
    keep, data, lens, paths, fa, n = _inputs(files, is_fasta)
    h = _lib().gsh_match_goal(db.h, meta.h, C.byref(c), data, lens, paths, fa, n,
                              filtered_path.encode() if filtered_path else None, kraken_path.encode() if kraken_path else None)
    return GoalResult(h)


def filter_goal(flt, k, files, is_fasta=None, min_pos_count=1, pos_ratio=0.2, with_probs=False, batch_reads=0, filtered_path=None,
                rest_path=None, want_rest=True, gpu_parse=True, text_chunk_bytes=0):
    keep, data, lens, paths, fa, n = _inputs(files, is_fasta)
    h = _lib().gsh_filter_goal(flt.h, k, min_pos_count, pos_ratio, int(with_probs), batch_reads, data, lens, paths, fa, n,
                               filtered_path.encode() if filtered_path else None, rest_path.encode() if rest_path else None, int(want_rest),
                               (int(text_chunk_bytes) if gpu_parse else 0xFFFFFFFF))
    return GoalResult(h)


def parse_only(k, files, is_fasta=None, with_probs=False):
    """Host parser alone (no GPU): .rest = every record rewritten by ReadEntry.write, totals, .accept = pooled entry index."""
    keep, data, lens, paths, fa, n = _inputs(files, is_fasta)
    return GoalResult(_lib().gsh_parse_only(k, int(with_probs), data, lens, paths, fa, n))


def last_record_start(text):
    """Where the GPU feeder would cut a text chunk: offset of the last record start (0 = no boundary found)."""
    b = np.frombuffer(text, dtype=np.uint8) if len(text) else np.zeros(1, dtype=np.uint8)
    return int(_lib().gsh_last_record_start(b.ctypes.data, len(text)))


def device_inflated_blocks():
    """Block-gzip members this process has inflated on the device so far (gs_inflate_blocks through the feeder)."""
    return int(_lib().gsh_device_inflated_blocks())


def bgzf_timers():
    """(seconds inside gs_inflate_blocks, calls, seconds inside the block-gzip reader) of this process so far."""
    out = (C.c_double * 3)()
    _lib().gsh_bgzf_timers(out)
    return float(out[0]), int(out[1]), float(out[2])


def bgzf_read_all(path, request=1 << 20):
    """The feeder's block-gzip reader alone (no GPU): (inflated text, compressed offset of a non-BGZF member that ended the
    fast path or -1), reading `request` bytes at a time."""
    foreign = C.c_int64(-1)
    res = GoalResult(_lib().gsh_bgzf_read_all(str(path).encode(), int(request), C.byref(foreign)))
    return res.rest, int(foreign.value)


def java_double_to_string(v):
    buf = C.create_string_buffer(64)
    n = _lib().gsh_java_double_to_string(v, buf, 64)
    return buf.value[:n].decode()
