"""ctypes binding of the C ABI in include/genestrip_b200.h (the same symbols the JNI shim binds).

Thin and typed: numpy arrays in, numpy arrays out; every error code becomes GenestripError with the
library's message.  There is no CPU fallback -- a missing extension or a missing CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

from .build import LIB_PATH

GS_ABI_VERSION = 8
GS_MAX_INFLIGHT = 3
GS_READ_FOUND, GS_READ_ACCEPTED, GS_READ_SLOWPATH = 1, 2, 4
GS_RUN_MISS, GS_RUN_INVALID = 0xFFFFFFFE, 0xFFFFFFFD
GS_BLOOM_BLOCKED, GS_BLOOM_XOR, GS_BLOOM_MURMUR = 0, 1, 2
GS_LAYOUT_TABLE, GS_LAYOUT_CLASSIC = 0, 1


class GenestripError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("genestrip_b200 error %d: %s" % (code, msg))
        self.code = code


class MatchCfg(C.Structure):
    """gs_match_cfg (include/genestrip_b200.h); defaults = C/GSConfigKey.java:302-350."""
    _fields_ = [("classify_reads", C.c_int), ("count_unique_kmers", C.c_int), ("max_kmer_res_counts", C.c_int),
                ("use_bloom_filter", C.c_int), ("max_classification_paths", C.c_int), ("min_kmers_for_class", C.c_int),
                ("max_read_tax_error_count", C.c_double), ("max_read_class_error_count", C.c_double),
                ("want_runs", C.c_int), ("layout", C.c_int), ("prefilter", C.c_int), ("host_pack_threads", C.c_int), ("host_pack_percent", C.c_int)]


class FastqInfo(C.Structure):
    """gs_fastq_info"""
    _fields_ = [("n_reads", C.c_uint32), ("status", C.c_uint32), ("total_kmers", C.c_uint64), ("total_bps", C.c_uint64)]


FASTQ_REC_DTYPE = np.dtype([("hdr_start", "<u4"), ("seq_start", "<u4"), ("seq_len", "<u4"), ("qual_start", "<u4")])
READ_RESULT_DTYPE = np.dtype([("class_vidx", "<i4"), ("read_kmers", "<u4"), ("tax_err", "<u4"), ("flags", "<u4")])
RUN_DTYPE = np.dtype([("label", "<u4"), ("len", "<u4")])
EVENT_DTYPE = np.dtype([("vidx", "<u4"), ("contig_len", "<u4"), ("read_no", "<u8")])
TAXON_COUNTS_DTYPE = np.dtype([("kmers", "<i8"), ("contigs", "<i8"), ("contig_len_squared_sum", "<i8"), ("reads_1kmer", "<i8"),
                               ("reads", "<i8"), ("reads_kmers", "<i8"), ("reads_bps", "<i8"), ("unique_kmers", "<i8"),
                               ("max_contig_len", "<i4"), ("touched", "<i4"), ("max_contig_read_no", "<u8")])
DEFLATE_BLOCK_DTYPE = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("in_len", "<u4"), ("out_len", "<u4"), ("crc32", "<u4"), ("status", "<u4")])
assert READ_RESULT_DTYPE.itemsize == 16 and TAXON_COUNTS_DTYPE.itemsize == 80 and EVENT_DTYPE.itemsize == 16 and DEFLATE_BLOCK_DTYPE.itemsize == 32

_lib = None

_P = C.c_void_p
_SIGS = {
    "gs_abi_version": (C.c_int, []),
    "gs_last_error": (C.c_char_p, []),
    "gs_ctx_create": (_P, [_P, C.c_int]),
    "gs_ctx_destroy": (None, [_P]),
    "gs_ctx_n_devices": (C.c_int, [_P]),
    "gs_alloc_pinned": (_P, [C.c_size_t]),
    "gs_free_pinned": (None, [_P]),
    "gs_db_create": (_P, [_P, C.c_int, C.c_uint64, C.c_int]),
    "gs_db_put_keys": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64]),
    "gs_db_put_values": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64]),
    "gs_db_put_radix_bucket": (C.c_int, [_P, C.c_int, C.c_uint32, _P, C.c_uint32]),
    "gs_db_set_tree": (C.c_int, [_P, _P, _P, C.c_int]),
    "gs_db_set_bloom_blocked": (C.c_int, [_P, C.c_int64, C.c_uint64, _P, C.c_uint64]),
    "gs_db_build_bloom_blocked": (C.c_int, [_P, _P, C.c_uint64]),
    "gs_db_finalize": (C.c_int, [_P]),
    "gs_db_update": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]),
    "gs_db_get_values": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64]),
    "gs_db_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "gs_db_save_file": (C.c_int, [_P, C.c_char_p]),
    "gs_db_load_file": (_P, [_P, C.c_char_p]),
    "gs_db_destroy": (None, [_P]),
    "gs_db_device_bytes": (C.c_uint64, [_P]),
    "gs_db_n_devices": (C.c_int, [_P]),
    "gs_db_lookup": (C.c_int, [_P, _P, C.c_uint64, C.c_int, _P, _P]),
    "gs_match_cfg_default": (None, [C.POINTER(MatchCfg)]),
    "gs_match_open": (_P, [_P, C.POINTER(MatchCfg)]),
    "gs_match_submit": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]),
    "gs_match_collect": (C.c_int, [_P, C.c_uint64, _P, _P, C.c_uint32, C.POINTER(C.c_uint32), _P, _P, C.c_uint64]),
    "gs_match_collect_view": (C.c_int, [_P, C.c_uint64, C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(_P), C.POINTER(C.c_uint32)]),
    "gs_match_submit_fastq": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.POINTER(FastqInfo), C.POINTER(C.c_uint64)]),
    "gs_match_collect_fastq": (C.c_int, [_P, C.c_uint64, C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(_P),
                                        _P, _P, C.c_uint64]),
    "gs_match_finish": (C.c_int, [_P, _P, _P]),
    "gs_comm_unique_id": (C.c_int, [_P]),
    "gs_comm_create": (_P, [_P, _P, C.c_int, C.c_int]),
    "gs_comm_world": (C.c_int, [_P]),
    "gs_comm_rank": (C.c_int, [_P]),
    "gs_comm_destroy": (None, [_P]),
    "gs_match_prepare_merge": (C.c_int, [_P, _P]),
    "gs_match_finish_comm": (C.c_int, [_P, _P, _P, _P]),
    "gs_pack_bases": (C.c_int, [_P, C.c_uint64, _P, _P, C.c_int]),
    "gs_pack_isa": (C.c_char_p, []),
    "gs_match_pack_fraction": (C.c_double, [_P]),
    "gs_match_l2_window": (C.c_int, [_P, _P, _P, _P]),
    "gs_match_join": (C.c_int, [_P]),
    "gs_match_pack_stats": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gs_match_merge_stats": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "gs_match_close": (None, [_P]),
    "gs_match_run_device": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint64, C.c_uint64, _P]),
    "gs_match_sync": (C.c_int, [_P]),
    "gs_match_device_state": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "gs_match_unique_popcount": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P]),
    "gs_match_stream": (_P, [_P]),
    "gs_match_kernel_launches": (C.c_uint64, [_P]),
    "gs_match_set_timing": (C.c_int, [_P, C.c_int]),
    "gs_match_kernel_times": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "gs_match_dump_labels": (C.c_int, [_P, _P, _P, C.c_uint32, _P, _P, _P]),
    "gs_filter_create": (_P, [_P, C.c_int, C.c_int64, C.c_int64, _P, _P, C.c_uint64]),
    "gs_filter_destroy": (None, [_P]),
    "gs_filter_n_devices": (C.c_int, [_P]),
    "gs_filter_contains": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "gs_filter_open": (_P, [_P, C.c_int, C.c_int, C.c_double]),
    "gs_filter_submit": (C.c_int, [_P, _P, _P, C.c_uint32, C.POINTER(C.c_uint64)]),
    "gs_filter_collect": (C.c_int, [_P, C.c_uint64, _P]),
    "gs_filter_submit_fastq": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(FastqInfo), C.POINTER(C.c_uint64)]),
    "gs_filter_collect_fastq": (C.c_int, [_P, C.c_uint64, C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(_P)]),
    "gs_filter_run_device": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint64, _P]),
    "gs_filter_kernel_launches": (C.c_uint64, [_P]),
    "gs_filter_save_file": (C.c_int, [_P, C.c_char_p]),
    "gs_filter_load_file": (_P, [_P, C.c_char_p]),
    "gs_filter_sync": (C.c_int, [_P]),
    "gs_filter_stream": (_P, [_P]),
    "gs_filter_close": (None, [_P]),
    "gs_inflate_blocks": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint32, _P, C.c_uint64]),
    "gs_db_context": (_P, [_P]),
    "gs_filter_context": (_P, [_P]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def lib():
    """Load the CUDA extension; fails loudly when it is missing (no CPU fallback)."""
    global _lib
    if _lib is None:
        path = os.environ.get("GS_LIB_VARIANT") or LIB_PATH  # tuning variants built by build_native(defines=..., out=...)
        if not os.path.exists(path):
            raise GenestripError(-2, "CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no CPU fallback)" % path)
        L = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.gs_abi_version() != GS_ABI_VERSION:
            raise GenestripError(-1, "ABI version mismatch")
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise GenestripError(rc, lib().gs_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def _arr(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class Context:
    """gs_ctx: the GPUs used by this process (replaces the consumer-thread pool, C/DefaultExecutionContext.java:78)."""

    def __init__(self, devices=None):
        L = lib()
        if devices:
            d = (C.c_int * len(devices))(*devices)
            self.h = L.gs_ctx_create(C.cast(d, _P), len(devices))
        else:
            self.h = L.gs_ctx_create(None, 0)
        if not self.h:
            raise GenestripError(-2, L.gs_last_error().decode())

    @property
    def n_devices(self):
        return lib().gs_ctx_n_devices(self.h)

    def inflate_blocks(self, comp, blocks, out_bytes):
        """gs_inflate_blocks: the raw-deflate members described by `blocks` (DEFLATE_BLOCK_DTYPE; see bgzf_blocks) of the
        block-gzip bytes `comp` -> text, inflated and checked (size, CRC-32) on the device.  Returns (text, blocks)."""
        comp = np.ascontiguousarray(np.frombuffer(comp, dtype=np.uint8))
        blocks = np.ascontiguousarray(blocks, dtype=DEFLATE_BLOCK_DTYPE).copy()
        out = np.empty(max(int(out_bytes), 1), dtype=np.uint8)
        rc = lib().gs_inflate_blocks(self.h, _ptr(comp), comp.size, _ptr(blocks), len(blocks), _ptr(out), int(out_bytes))
        self.last_inflate_blocks = blocks
        _check(rc)
        return out[:int(out_bytes)], blocks

    def close(self):
        if self.h:
            lib().gs_ctx_destroy(self.h)
            self.h = None


class Comm:
    """gs_comm: this process' GPU as rank `rank` of `world` (one process per GPU); NCCL underneath, loaded at run time."""

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        _check(lib().gs_comm_unique_id(C.cast(buf, _P)))
        return bytes(buf)

    def __init__(self, ctx, unique_id, world, rank):
        assert len(unique_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self.h = lib().gs_comm_create(ctx.h, C.cast(buf, _P), int(world), int(rank))
        if not self.h:
            raise GenestripError(-2, lib().gs_last_error().decode())
        self.world, self.rank = int(world), int(rank)

    def close(self):
        if self.h:
            lib().gs_comm_destroy(self.h)
            self.h = None


def pack_bases(bases, threads=1):
    """gs_pack_bases: ASCII bases -> (codes uint64[ceil(n/32)], valid uint32[ceil(n/32)]), the form host_pack_threads puts on the link."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    words = (len(bases) + 31) // 32
    codes = np.empty(words, dtype=np.uint64)
    valid = np.empty(words, dtype=np.uint32)
    _check(lib().gs_pack_bases(_ptr(bases), len(bases), _ptr(codes), _ptr(valid), int(threads)))
    return codes, valid


def bgzf_blocks(comp):
    """Member table of a block-gzip (BGZF) byte string for Context.inflate_blocks: one entry per member, from the 'BC' extra
    subfield (member size - 1) and the trailer (CRC-32, ISIZE).  Returns (blocks, total inflated bytes)."""
    import struct
    rows, off, out = [], 0, 0
    while off < len(comp):
        if comp[off:off + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a block-gzip member at byte %d" % off)
        xlen = struct.unpack_from("<H", comp, off + 10)[0]
        p, total = off + 12, None
        while p + 4 <= off + 12 + xlen:
            si1, si2, slen = struct.unpack_from("<BBH", comp, p)
            if (si1, si2, slen) == (66, 67, 2):
                total = struct.unpack_from("<H", comp, p + 4)[0] + 1
            p += 4 + slen
        if total is None:
            raise ValueError("member at byte %d has no BC subfield" % off)
        crc, isize = struct.unpack_from("<II", comp, off + total - 8)
        rows.append((off + 12 + xlen, out, total - 12 - xlen - 8, isize, crc, 0))
        off += total
        out += isize
    return np.array(rows, dtype=DEFLATE_BLOCK_DTYPE), out


class PinnedBuffer:
    """gs_alloc_pinned as a numpy uint8 view (the pinned, double-buffered host batches of the parser)."""

    def __init__(self, nbytes):
        self.ptr = lib().gs_alloc_pinned(nbytes)
        if not self.ptr:
            raise GenestripError(-2, lib().gs_last_error().decode())
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def view(self, dtype, count, offset=0):
        return self.array[offset:offset + count * np.dtype(dtype).itemsize].view(dtype)

    def free(self):
        if self.ptr:
            self.array = None
            lib().gs_free_pinned(self.ptr)
            self.ptr = None


class Database:
    """gs_db: sorted k-mer store + value indices + tax tree + blocked Bloom prefilter, device resident."""

    def __init__(self, ctx, k, keys, vidx_raw, n_values, parent_by_vidx=None, has_node=None, bloom=None, build_bloom=False,
                 radix=None):
        L = lib()
        self.ctx = ctx
        self.k = k
        self.n_values = int(n_values)
        if radix is None:
            keys = _arr(keys, np.int64)
            vidx_raw = _arr(vidx_raw, np.int16)
            self.n_kmers = int(keys.shape[0])
        else:
            self.n_kmers = int(sum(len(e) for _, e in radix[1]))
        self.h = L.gs_db_create(ctx.h, k, self.n_kmers, self.n_values)
        if not self.h:
            raise GenestripError(-1, L.gs_last_error().decode())
        try:
            if radix is None:
                seg = 1 << 24  # streamed in segments like the reference's BigArrays (2^27) -- smaller here to exercise offsets
                for off in range(0, self.n_kmers, seg):
                    n = min(seg, self.n_kmers - off)
                    _check(L.gs_db_put_keys(self.h, off, _ptr(keys[off:off + n]), n))
                    _check(L.gs_db_put_values(self.h, off, _ptr(vidx_raw[off:off + n]), n))
            else:
                radix_bits, buckets = radix
                for r, entries in buckets:
                    e = _arr(entries, np.int64)
                    _check(L.gs_db_put_radix_bucket(self.h, radix_bits, int(r), _ptr(e), len(e)))
            if parent_by_vidx is not None:
                p = _arr(parent_by_vidx, np.int32)
                hn = None if has_node is None else _arr(has_node, np.int32)
                _check(L.gs_db_set_tree(self.h, _ptr(p), _ptr(hn), self.n_values))
            self.bloom_words = None
            if bloom is not None:
                seed, buckets, words = bloom
                words = _arr(words, np.int64)
                _check(L.gs_db_set_bloom_blocked(self.h, int(seed), int(buckets), _ptr(words), len(words)))
            elif build_bloom:
                nb = (max(1, self.n_kmers) * 10 + 63) // 64
                out = np.zeros(nb + 17, dtype=np.int64) if build_bloom == "return" else None
                _check(L.gs_db_build_bloom_blocked(self.h, _ptr(out), 0 if out is None else len(out)))
                self.bloom_words = out
            _check(L.gs_db_finalize(self.h))
        except Exception:
            L.gs_db_destroy(self.h)
            self.h = None
            raise

    @classmethod
    def from_pointers(cls, ctx, k, keys_ptr, vals_ptr, n_kmers, n_values, parent_by_vidx, build_bloom=True):
        """Database whose sorted keys / raw values already sit in (device or host) memory at the given addresses."""
        L = lib()
        self = cls.__new__(cls)
        self.ctx, self.k, self.n_values, self.n_kmers, self.bloom_words = ctx, k, int(n_values), int(n_kmers), None
        self.h = L.gs_db_create(ctx.h, k, self.n_kmers, self.n_values)
        if not self.h:
            raise GenestripError(-1, L.gs_last_error().decode())
        try:
            _check(L.gs_db_put_keys(self.h, 0, keys_ptr, self.n_kmers))
            _check(L.gs_db_put_values(self.h, 0, vals_ptr, self.n_kmers))
            p = _arr(parent_by_vidx, np.int32)
            _check(L.gs_db_set_tree(self.h, _ptr(p), None, self.n_values))
            if build_bloom:
                _check(L.gs_db_build_bloom_blocked(self.h, None, 0))
            _check(L.gs_db_finalize(self.h))
        except Exception:
            L.gs_db_destroy(self.h)
            self.h = None
            raise
        return self

    @property
    def device_bytes(self):
        return lib().gs_db_device_bytes(self.h)

    def save(self, path):
        """Flat GSB1 file (keys, Java-short values, tree by value index, blocked Bloom filter)."""
        _check(lib().gs_db_save_file(self.h, str(path).encode()))

    @classmethod
    def load(cls, ctx, path):
        """Database from a GSB1 file, without the arrays passing through Python."""
        L = lib()
        self = cls.__new__(cls)
        self.ctx = ctx
        self.h = L.gs_db_load_file(ctx.h, str(path).encode())
        if not self.h:
            raise GenestripError(-1, L.gs_last_error().decode())
        k, n, v = C.c_int(0), C.c_uint64(0), C.c_int(0)
        _check(L.gs_db_info(self.h, C.byref(k), C.byref(n), C.byref(v)))
        self.k, self.n_kmers, self.n_values, self.bloom_words = k.value, int(n.value), v.value, None
        return self

    def update(self, seq, region_offsets, region_vidx, upper_case=True):
        """DBGoal update phase: value = LCA(value, region node) for every stored k-mer of the regions; returns #changes."""
        seq = _arr(seq, np.uint8)
        off = _arr(region_offsets, np.uint64)
        vid = _arr(region_vidx, np.int32)
        ch = C.c_uint64(0)
        _check(lib().gs_db_update(self.h, _ptr(seq), len(seq), _ptr(off), _ptr(vid), len(vid), int(upper_case), C.byref(ch)))
        return ch.value

    def values(self, offset=0, n=None):
        n = self.n_kmers - offset if n is None else n
        out = np.empty(n, dtype=np.int16)
        _check(lib().gs_db_get_values(self.h, offset, _ptr(out), n))
        return out

    def lookup(self, kmers, use_bloom=True):
        kmers = _arr(kmers, np.int64)
        v = np.empty(len(kmers), dtype=np.int32)
        p = np.empty(len(kmers), dtype=np.int64)
        _check(lib().gs_db_lookup(self.h, _ptr(kmers), len(kmers), 1 if use_bloom else 0, _ptr(v), _ptr(p)))
        return v, p

    def close(self):
        if self.h:
            lib().gs_db_destroy(self.h)
            self.h = None


def default_match_cfg(**over):
    cfg = MatchCfg()
    lib().gs_match_cfg_default(C.byref(cfg))
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


class MatchSession:
    """gs_sess: one FastqKMerMatcher.runMatcher run (C/match/FastqKMerMatcher.java:181-235)."""

    def __init__(self, db, cfg=None):
        self.db = db
        self.cfg = cfg if cfg is not None else default_match_cfg()
        self.h = lib().gs_match_open(db.h, C.byref(self.cfg))
        if not self.h:
            raise GenestripError(-1, lib().gs_last_error().decode())
        self._keep = {}

    def submit(self, bases, offsets, first_read_no):
        """bases: uint8 array (host, ideally pinned); offsets: uint64[n+1].  Returns a ticket."""
        n = len(offsets) - 1
        t = C.c_uint64(0)
        _check(lib().gs_match_submit(self.h, _ptr(bases), _ptr(offsets), n, int(first_read_no), C.byref(t)))
        self._keep[t.value] = (bases, offsets, n)
        return t.value

    def collect(self, ticket, want_events=True):
        bases, offsets, n = self._keep.pop(ticket)
        out = np.empty(n, dtype=READ_RESULT_DTYPE)
        ev = np.empty(max(self.db.n_values, 1), dtype=EVENT_DTYPE)
        nev = C.c_uint32(0)
        runs = run_off = None
        if self.cfg.want_runs:
            k = self.db.k
            lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
            cap = int(np.maximum(lens - k + 1, 0).sum())
            run_off = np.zeros(n + 1, dtype=np.uint64)
            runs = np.empty(max(cap, 1), dtype=RUN_DTYPE)
            _check(lib().gs_match_collect(self.h, ticket, _ptr(out), _ptr(ev), len(ev), C.byref(nev), _ptr(run_off), _ptr(runs), cap))
            runs = runs[:int(run_off[n])]
        else:
            _check(lib().gs_match_collect(self.h, ticket, _ptr(out), _ptr(ev), len(ev), C.byref(nev), None, None, 0))
        return out, ev[:nev.value].copy(), run_off, runs

    def collect_view(self, ticket):
        """Zero-copy collect: numpy views of the session's pinned result / event staging (valid until the slot is reused)."""
        self._keep.pop(ticket)
        out, ev = _P(), _P()
        n, nev = C.c_uint32(0), C.c_uint32(0)
        _check(lib().gs_match_collect_view(self.h, ticket, C.byref(out), C.byref(n), C.byref(ev), C.byref(nev)))
        res = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(n.value * 16,)).view(READ_RESULT_DTYPE) if n.value else np.zeros(0, READ_RESULT_DTYPE)
        evs = np.ctypeslib.as_array(C.cast(ev, C.POINTER(C.c_uint8)), shape=(nev.value * 16,)).view(EVENT_DTYPE) if nev.value else np.zeros(0, EVENT_DTYPE)
        return res, evs

    def submit_fastq(self, text, first_read_no=0, n_bytes=None):
        """Raw FASTQ text (whole records, '\\n'-terminated).  Returns (ticket, FastqInfo); ticket == 0 means the chunk is not
        strict 4-line FASTQ (info.status holds the GS_FASTQ_* bits) and must be parsed on the CPU."""
        text = _arr(text, np.uint8)
        n = len(text) if n_bytes is None else int(n_bytes)
        info = FastqInfo()
        t = C.c_uint64(0)
        _check(lib().gs_match_submit_fastq(self.h, _ptr(text), n, int(first_read_no), C.byref(info), C.byref(t)))
        if t.value:
            self._keep[t.value] = (text, int(info.total_kmers), info.n_reads)
        return t.value, info

    def collect_fastq(self, ticket):
        """(results, events, event header offsets, records[n + 1]) as numpy views of the session's pinned staging; with
        want_runs two more items: run_offsets[n + 1], runs."""
        _, total_kmers, n_reads = self._keep.pop(ticket)
        out, ev, eh, rc = _P(), _P(), _P(), _P()
        n, nev = C.c_uint32(0), C.c_uint32(0)
        run_off = runs = None
        if self.cfg.want_runs:
            run_off = np.zeros(n_reads + 1, dtype=np.uint64)
            runs = np.empty(max(total_kmers, 1), dtype=RUN_DTYPE)
        _check(lib().gs_match_collect_fastq(self.h, ticket, C.byref(out), C.byref(n), C.byref(ev), C.byref(eh), C.byref(nev), C.byref(rc),
                                            _ptr(run_off), _ptr(runs), total_kmers if runs is not None else 0))
        view = lambda p, cnt, size, dt: (np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(cnt * size,)).view(dt) if cnt else np.zeros(0, dt))
        res = (view(out, n.value, 16, READ_RESULT_DTYPE), view(ev, nev.value, 16, EVENT_DTYPE), view(eh, nev.value, 4, np.dtype("<u4")),
               view(rc, n.value + 1, 16, FASTQ_REC_DTYPE))
        if runs is not None:
            res = res + (run_off, runs[:int(run_off[n_reads])])
        return res

    def prepare_merge(self, comm):
        """gs_match_prepare_merge (collective): peer mappings and merge kernel set up ahead of the run."""
        _check(lib().gs_match_prepare_merge(self.h, comm.h))

    def finish(self, comm=None):
        """End of the run.  With `comm` (one process per GPU): the collective merge over all ranks, gs_match_finish_comm."""
        V = self.db.n_values
        counts = np.zeros(max(V, 1), dtype=TAXON_COUNTS_DTYPE)
        top = None
        if self.cfg.count_unique_kmers and self.cfg.max_kmer_res_counts > 0:
            top = np.zeros((V + 1, self.cfg.max_kmer_res_counts), dtype=np.int16)
        if comm is None:
            _check(lib().gs_match_finish(self.h, _ptr(counts), _ptr(top)))
        else:
            _check(lib().gs_match_finish_comm(self.h, comm.h, _ptr(counts), _ptr(top)))
        return counts[:V], top

    @property
    def pack_fraction(self):
        return lib().gs_match_pack_fraction(self.h)

    def join(self):
        """Orders the compute stream behind everything submitted so far (no host wait): before an event of one's own."""
        _check(lib().gs_match_join(self.h))

    def l2_window(self):
        """(window bytes, persisting carve-out bytes, hit ratio) of the L2 access policy window on the compute stream."""
        w, c, r = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
        _check(lib().gs_match_l2_window(self.h, C.byref(w), C.byref(c), C.byref(r)))
        return int(w.value), int(c.value), float(r.value)

    def pack_stats(self):
        """(threads, host seconds spent packing, bases packed, base bytes put on the link) of this session's submits."""
        t, sec, n, b = C.c_int(0), C.c_double(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().gs_match_pack_stats(self.h, C.byref(t), C.byref(sec), C.byref(n), C.byref(b)))
        return t.value, sec.value, n.value, b.value

    def merge_stats(self):
        """(total_ms, bitset_ms, bytes read from the other ranks, path) of the last multi-GPU merge; path 1 = peer mappings, 2 = NCCL."""
        a, b, n, p = C.c_double(0), C.c_double(0), C.c_uint64(0), C.c_int(0)
        _check(lib().gs_match_merge_stats(self.h, C.byref(a), C.byref(b), C.byref(n), C.byref(p)))
        return a.value, b.value, n.value, p.value

    # ---- device-resident variants (bench kernel-only number, NCCL reduction of raw state)
    def run_device(self, d_bases_ptr, d_offsets_ptr, n_reads, n_bases, first_read_no, d_out_ptr):
        _check(lib().gs_match_run_device(self.h, d_bases_ptr, d_offsets_ptr, n_reads, int(n_bases), int(first_read_no), d_out_ptr))

    def sync(self):
        _check(lib().gs_match_sync(self.h))

    def device_state(self):
        c, m, b = _P(), _P(), _P()
        w = C.c_uint64(0)
        _check(lib().gs_match_device_state(self.h, C.byref(c), C.byref(m), C.byref(b), C.byref(w)))
        return c.value, m.value, b.value, w.value

    def unique_popcount(self, d_bitset_ptr, word_begin, word_end, d_unique_ptr):
        _check(lib().gs_match_unique_popcount(self.h, d_bitset_ptr, word_begin, word_end, d_unique_ptr))

    def dump_labels(self, d_bases_ptr, d_offsets_ptr, n_reads, d_kmer_offsets_ptr, d_labels_ptr, d_pos_ptr):
        _check(lib().gs_match_dump_labels(self.h, d_bases_ptr, d_offsets_ptr, n_reads, d_kmer_offsets_ptr, d_labels_ptr, d_pos_ptr))

    @property
    def stream(self):
        return lib().gs_match_stream(self.h)

    @property
    def kernel_launches(self):
        return lib().gs_match_kernel_launches(self.h)

    def set_timing(self, on=True):
        _check(lib().gs_match_set_timing(self.h, 1 if on else 0))

    def kernel_times(self):
        """(label_ms, reduce_ms, n_batches): average CUDA-event durations per batch since set_timing(True)."""
        a, b, n = C.c_double(0), C.c_double(0), C.c_uint64(0)
        _check(lib().gs_match_kernel_times(self.h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def close(self):
        if self.h:
            lib().gs_match_close(self.h)
            self.h = None


class Filter:
    """gs_filter: the `filter` goal's KMerProbFilter index, device resident (C/goals/LoadIndexGoal.java:92-104)."""

    def __init__(self, ctx, kind, p0, p1, factors, words):
        words = _arr(words, np.int64)
        f = None if factors is None else _arr(factors, np.int64)
        self.h = lib().gs_filter_create(ctx.h, kind, int(p0), int(p1), _ptr(f), _ptr(words), len(words))
        if not self.h:
            raise GenestripError(-1, lib().gs_last_error().decode())

    def contains(self, kmers):
        kmers = _arr(kmers, np.int64)
        out = np.empty(len(kmers), dtype=np.uint8)
        _check(lib().gs_filter_contains(self.h, _ptr(kmers), len(kmers), _ptr(out)))
        return out

    def save_file(self, path):
        """Flat GSF1 index file (gs_filter_save_file)."""
        _check(lib().gs_filter_save_file(self.h, os.fsencode(path)))

    @classmethod
    def load_file(cls, ctx, path):
        """gs_filter_load_file: the index KMerProbFilter.load would deserialize, without a JVM."""
        self = cls.__new__(cls)
        self.h = lib().gs_filter_load_file(ctx.h, os.fsencode(path))
        if not self.h:
            raise GenestripError(-1, lib().gs_last_error().decode())
        return self

    def close(self):
        if self.h:
            lib().gs_filter_destroy(self.h)
            self.h = None


class FilterSession:
    """gs_fsess: one FastqBloomFilter.runFilter run (C/bloom/FastqBloomFilter.java:80-161)."""

    def __init__(self, flt, k, min_pos_count=1, pos_ratio=0.2):
        self.h = lib().gs_filter_open(flt.h, k, min_pos_count, pos_ratio)
        if not self.h:
            raise GenestripError(-1, lib().gs_last_error().decode())
        self._keep = {}

    def submit(self, bases, offsets):
        n = len(offsets) - 1
        t = C.c_uint64(0)
        _check(lib().gs_filter_submit(self.h, _ptr(bases), _ptr(offsets), n, C.byref(t)))
        self._keep[t.value] = (bases, offsets, n)
        return t.value

    def collect(self, ticket):
        _, _, n = self._keep.pop(ticket)
        out = np.empty(n, dtype=np.uint8)
        _check(lib().gs_filter_collect(self.h, ticket, _ptr(out)))
        return out

    def submit_fastq(self, text, n_bytes=None):
        """Raw FASTQ text; returns (ticket, FastqInfo), ticket == 0 if the chunk is not strict 4-line FASTQ."""
        text = _arr(text, np.uint8)
        n = len(text) if n_bytes is None else int(n_bytes)
        info = FastqInfo()
        t = C.c_uint64(0)
        _check(lib().gs_filter_submit_fastq(self.h, _ptr(text), n, C.byref(info), C.byref(t)))
        if t.value:
            self._keep[t.value] = (text, None, info.n_reads)
        return t.value, info

    def collect_fastq(self, ticket):
        self._keep.pop(ticket)
        acc, rc = _P(), _P()
        n = C.c_uint32(0)
        _check(lib().gs_filter_collect_fastq(self.h, ticket, C.byref(acc), C.byref(n), C.byref(rc)))
        a = np.ctypeslib.as_array(C.cast(acc, C.POINTER(C.c_uint8)), shape=(n.value,)) if n.value else np.zeros(0, np.uint8)
        r = np.ctypeslib.as_array(C.cast(rc, C.POINTER(C.c_uint8)), shape=((n.value + 1) * 16,)).view(FASTQ_REC_DTYPE)
        return a, r

    def run_device(self, d_bases_ptr, d_offsets_ptr, n_reads, n_bases, d_accept_ptr):
        _check(lib().gs_filter_run_device(self.h, d_bases_ptr, d_offsets_ptr, n_reads, n_bases, d_accept_ptr))

    @property
    def kernel_launches(self):
        return int(lib().gs_filter_kernel_launches(self.h))

    def sync(self):
        _check(lib().gs_filter_sync(self.h))

    @property
    def stream(self):
        return lib().gs_filter_stream(self.h)

    def close(self):
        if self.h:
            lib().gs_filter_close(self.h)
            self.h = None
