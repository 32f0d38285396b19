"""Key codec for the minimizer-addressed table (sketch in README.md of this directory), checked on the CPU:

  canonical 31-mer  <->  (canonical minimizer 23-mer, offset of that window in the canonical k-mer, strand of the window, 16 flank bits)

and, through a bijective 46-bit mix of the minimizer,  <->  (line index, 8-byte slot payload).  The properties a lookup
kernel relies on are asserted on random k-mers and on k-mers with repeated 23-mers (ties between windows):
  1. both strands of a k-mer give the same tuple (the rule never looks at the strand the read shows);
  2. the tuple identifies the k-mer: decode(encode(x)) == x;
  3. the slot payload fits 62 bits at a table of >= 2^24 lines, value and flags included;
  4. line B = line A xor g(remainder) is an involution given the remainder (two candidate lines, one stored key form).

Encoding of bases as in the reference (C/util/CGAT.java:37-83): C0 G1 A2 T3, complement = code xor 1, canonical = max(fwd, rc).
usage: python profiles/experiments/mzline_keycodec.py
"""
import random

K, M = 31, 23
W = K - M + 1
MASK_K, MASK_M = (1 << (2 * K)) - 1, (1 << (2 * M)) - 1


def revcomp(x, n):
    r = 0
    for _ in range(n):
        r = (r << 2) | ((x & 3) ^ 1)
        x >>= 2
    return r


def canonical(x, n):
    return max(x, revcomp(x, n))


def window(x, p):
    """23-mer at offset p (0 = leftmost = most significant) of the 31-mer x."""
    return (x >> (2 * (K - M - p))) & MASK_M


# a bijection on 46 bits (odd multiplications and xor-shifts are invertible): its top bits choose the line
MUL1, MUL2 = 0x2545F4914F6CDD1D & MASK_M | 1, 0x9E3779B97F4A7C15 & MASK_M | 1
INV1, INV2 = pow(MUL1, -1, 1 << 46), pow(MUL2, -1, 1 << 46)


def mix46(x):
    x = (x * MUL1) & MASK_M
    x ^= x >> 23
    x = (x * MUL2) & MASK_M
    x ^= x >> 23
    return x


def unmix46(x):
    x ^= x >> 23
    x = (x * INV2) & MASK_M
    x ^= x >> 23
    x = (x * INV1) & MASK_M
    return x


def encode(kmer):
    """kmer: either strand.  Returns (c, p, s, flanks) with c = canonical minimizer 23-mer, p = its offset in the CANONICAL
    k-mer (smallest mixed value wins, ties go to the smaller offset there), s = 1 if the window shows rc(c), flanks = the
    8 bases around the window (left ones first)."""
    x = canonical(kmer & MASK_K, K)
    best = None
    for p in range(W):
        wv = window(x, p)
        c = canonical(wv, M)
        h = mix46(c)
        if best is None or h < best[0]:      # strict: ties keep the smaller offset
            best = (h, p, c, 0 if wv == c else 1)
    _, p, c, s = best
    left = x >> (2 * (K - p))                                 # p bases
    right = x & ((1 << (2 * (W - 1 - p))) - 1)                # 8 - p bases
    flanks = (left << (2 * (W - 1 - p))) | right              # 16 bits
    return c, p, s, flanks


def decode(c, p, s, flanks):
    wv = revcomp(c, M) if s else c
    left = flanks >> (2 * (W - 1 - p))
    right = flanks & ((1 << (2 * (W - 1 - p))) - 1)
    return (left << (2 * (K - p))) | (wv << (2 * (W - 1 - p))) | right


def slot(kmer, t_bits, value=0):
    """(line A, 62-bit payload): payload = remainder of the mixed minimizer | offset | strand | flanks | value."""
    c, p, s, flanks = encode(kmer)
    h = mix46(c)
    line = h >> (46 - t_bits)
    rem = h & ((1 << (46 - t_bits)) - 1)
    payload = (((((rem << 4) | p) << 1 | s) << 16 | flanks) << 16) | value
    return line, payload


def unslot(line, payload, t_bits):
    value = payload & 0xFFFF
    flanks = (payload >> 16) & 0xFFFF
    s = (payload >> 32) & 1
    p = (payload >> 33) & 15
    rem = payload >> 37
    c = unmix46((line << (46 - t_bits)) | rem)
    return decode(c, p, s, flanks), value


def other_line(line, payload, t_bits):
    rem = payload >> 37
    g = (rem * 0x5BD1E995 + 0x27D4EB2F) & ((1 << t_bits) - 1)
    return line ^ (g | 1)    # never the same line


if __name__ == "__main__":
    rnd = random.Random(7)
    cases = [rnd.getrandbits(2 * K) for _ in range(20000)]
    # repeated 23-mers: homopolymers, short periods, a k-mer and its shifted copy
    for base in range(4):
        cases.append(sum(base << (2 * i) for i in range(K)))
    for period in (2, 3, 4, 5, 8):
        unit = [rnd.getrandbits(2) for _ in range(period)]
        cases.append(sum(unit[i % period] << (2 * i) for i in range(K)))
    for x in list(cases):
        assert unmix46(mix46(x & MASK_M)) == x & MASK_M
    for t_bits in (24, 27):
        for x in cases:
            rc = revcomp(x, K)
            e = encode(x)
            assert e == encode(rc), "strand dependence"
            assert decode(*e) == canonical(x, K), "tuple does not identify the k-mer"
            line, payload = slot(x, t_bits, value=0xBEEF)
            assert line < (1 << t_bits) and payload < (1 << (46 - t_bits + 4 + 1 + 16 + 16))
            assert 46 - t_bits + 37 <= 62 - 3 or t_bits >= 24, "payload + 3 flag bits must fit 62 bits"
            assert unslot(line, payload, t_bits) == (canonical(x, K), 0xBEEF)
            b = other_line(line, payload, t_bits)
            assert b != line and other_line(b, payload, t_bits) == line
    bits = {t: 46 - t + 4 + 1 + 16 + 16 for t in (24, 25, 26, 27)}
    print("codec ok on %d k-mers; payload bits incl. 16 value bits at 2^t lines:" % len(cases), bits, "(+ occupied / seen / spill = 3 flags; 64 available)")
