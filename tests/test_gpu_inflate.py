"""Block-gzip (BGZF) members inflated on the device (gs_inflate_blocks, genestrip_b200/csrc/gs_inflate.cu) against zlib:
the text java.util.zip.GZIPInputStream / gzread would hand to the FASTQ reader (C/fastq/AbstractFastqReader.java:224)."""
import gzip
import zlib

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _fastq(rng, n):
    recs = []
    for i in range(n):
        ln = int(rng.integers(30, 300))
        seq = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=ln, p=[0.24, 0.25, 0.25, 0.25, 0.01]).tobytes()
        recs.append(b"@r%d x\n" % i + seq + b"\n+\n" + bytes(rng.integers(33, 74, size=ln, dtype=np.uint8)) + b"\n")
    return b"".join(recs)


def test_inflate_blocks_matches_zlib(native, gpu_ctx):
    """Stored, fixed and dynamic deflate blocks, all compression levels, block sizes from 37 bytes to 64 KB, empty members."""
    rng = np.random.default_rng(1)
    cases = {"fastq": _fastq(rng, 6000), "random": bytes(rng.integers(0, 256, size=200000, dtype=np.uint8)), "zeros": b"\0" * 300000,
             "short": b"A", "empty": b"", "text": b"the quick brown fox " * 5000}
    for name, data in cases.items():
        for level in (0, 1, 6, 9):
            for block in (0xff00, 1000, 37):
                d = data[:20000] if block < 1000 and len(data) > 50000 else data
                comp = util.bgzf_bytes(d, block=block, level=level)
                assert gzip.decompress(comp) == d
                blocks, n = native.bgzf_blocks(comp)
                out, st = gpu_ctx.inflate_blocks(comp, blocks, n)
                assert n == len(d) and out.tobytes() == d and not st["status"].any(), (name, level, block)
    # fixed Huffman codes (zlib never picks them for blocks this long on its own)
    d = cases["fastq"][:60000]
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
    body = co.compress(d) + co.flush()
    blocks = np.array([(0, 0, len(body), len(d), zlib.crc32(d), 0)], dtype=native.DEFLATE_BLOCK_DTYPE)
    out, _ = gpu_ctx.inflate_blocks(body, blocks, len(d))
    assert out.tobytes() == d


def test_inflate_blocks_detects_corruption(native, gpu_ctx):
    """A flipped bit anywhere in the deflate data or the trailer is an error (GS_ERR_DATA), as gzread's data error /
    incorrect data check and GZIPInputStream's ZipException; the blocks that are fine are marked fine; and no input makes
    the kernel run away (every loop is bounded by the block's input bits and output bytes)."""
    rng = np.random.default_rng(2)
    data = _fastq(rng, 1500)
    comp = util.bgzf_bytes(data)
    blocks, n = native.bgzf_blocks(comp)
    for trial in range(40):
        c = bytearray(comp)
        b = int(rng.integers(0, len(blocks) - 1))                      # not the empty end-of-file block
        pos = int(blocks["in_off"][b]) + int(rng.integers(0, int(blocks["in_len"][b])))
        c[pos] ^= 1 << int(rng.integers(0, 8))
        with pytest.raises(native.GenestripError) as e:
            gpu_ctx.inflate_blocks(bytes(c), blocks, n)
        assert e.value.code == -5
        st = gpu_ctx.last_inflate_blocks["status"]
        assert st[b] != 0 and not np.delete(st, b).any()
    for field, delta in (("in_len", -5), ("out_len", -5), ("crc32", 1)):
        b2 = blocks.copy()
        b2[field][0] = int(b2[field][0]) + delta
        with pytest.raises(native.GenestripError):
            gpu_ctx.inflate_blocks(comp, b2, n)
    with pytest.raises(native.GenestripError) as e:                    # a block outside the buffers is refused up front
        b2 = blocks.copy()
        b2["in_off"][0] = len(comp)
        gpu_ctx.inflate_blocks(comp, b2, n)
    assert e.value.code == -1
    # output ranges that overlap are refused (the blocks are inflated concurrently); ranges that leave gaps are accepted and
    # the gaps come back as zero bytes, whatever an earlier call left in the device buffer
    if len(blocks) >= 3:
        b2 = blocks.copy()
        b2["out_off"][1] = int(b2["out_off"][1]) - 1
        with pytest.raises(native.GenestripError) as e:
            gpu_ctx.inflate_blocks(comp, b2, n)
        assert e.value.code == -1
        gpu_ctx.inflate_blocks(comp, blocks, n)                         # fills the device buffer with text
        b3 = blocks.copy()
        shift = 64
        b3["out_off"][1:] = b3["out_off"][1:] + shift                   # a hole of 64 bytes behind the first block
        out, _ = gpu_ctx.inflate_blocks(comp, b3, n + shift)
        e0 = int(blocks["out_off"][1])
        assert out[:e0].tobytes() == data[:e0] and not out[e0:e0 + shift].any() and out[e0 + shift:].tobytes() == data[e0:]
