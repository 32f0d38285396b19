"""Seeded synthetic projects of the shapes BASELINE.json names (genomes, nodes.dmp/names.dmp, reads).

Pure numpy; used by tests, smoke() and bench.py to make inputs for BOTH the CUDA path and the oracle.
Nothing here computes results.
"""
import numpy as np

BASES = np.frombuffer(b"CGAT", dtype=np.uint8)  # code order of C/util/CGAT.java:66-69
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in ((b"A", b"T"), (b"C", b"G"), (b"G", b"C"), (b"T", b"A"), (b"N", b"N")):
    _COMP[_a[0]] = _b[0]


def random_genome(rng, n):
    return BASES[rng.integers(0, 4, size=n)]


def revcomp(seq):
    return _COMP[seq[::-1]]


def tiny_taxonomy():
    """root 1; genera 10, 20; species 11, 12, 13 under 10 and 21, 22 under 20."""
    edges = [(1, 1, "no rank"), (10, 1, "genus"), (20, 1, "genus"), (11, 10, "species"), (12, 10, "species"),
             (13, 10, "species"), (21, 20, "species"), (22, 20, "species")]
    nodes = "".join("%d\t|\t%d\t|\t%s\t|\t\t|\n" % e for e in edges)
    names = "".join("%d\t|\ttaxon %d\t|\t\t|\tscientific name\t|\n" % (e[0], e[0]) for e in edges)
    return nodes, names, [11, 12, 13, 21, 22]


def tiny_project(genome_len=1_000_000, seed=42, shared_frac=0.005):
    """C1: 5 genomes, 0.5 % of each copied from the next sibling so that LCA != leaf occurs."""
    nodes, names, leaves = tiny_taxonomy()
    rngs = [np.random.default_rng(seed + g) for g in range(len(leaves))]
    seqs = [random_genome(r, genome_len) for r in rngs]
    n_shared = int(genome_len * shared_frac)
    for g in range(len(leaves)):
        src = seqs[(g + 1) % len(leaves)]
        a = int(rngs[g].integers(0, genome_len - n_shared))
        b = int(rngs[g].integers(0, genome_len - n_shared))
        seqs[g][a:a + n_shared] = src[b:b + n_shared]
    # a stretch shared between the two genera -> LCA = root
    seqs[4][1000:1000 + 500] = seqs[0][5000:5500]
    genomes = [(str(t), s.tobytes()) for t, s in zip(leaves, seqs)]
    return nodes, names, genomes


def sample_reads(genomes, n_reads, read_len, seed, frac_db=0.7, sub_rate=0.01, n_rate=0.001, len_jitter=0):
    """Illumina-like reads: frac_db sampled from the genomes (random strand, substitutions, N), rest iid random.

    Returns (bases uint8[total], offsets uint64[n+1], src int32[n]) with src = genome index or -1.
    """
    rng = np.random.default_rng(seed)
    glens = np.array([len(g) for g in genomes], dtype=np.int64)
    gcat = np.concatenate([np.frombuffer(g, dtype=np.uint8) if isinstance(g, (bytes, bytearray)) else g for g in genomes])
    gstart = np.concatenate([[0], np.cumsum(glens)[:-1]])
    if len_jitter:
        lens = rng.integers(max(1, read_len - len_jitter), read_len + len_jitter + 1, size=n_reads).astype(np.int64)
    else:
        lens = np.full(n_reads, read_len, dtype=np.int64)
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens).astype(np.uint64)
    total = int(offsets[-1])
    from_db = rng.random(n_reads) < frac_db
    gi = rng.integers(0, len(genomes), size=n_reads)
    maxstart = np.maximum(glens[gi] - lens, 0)
    st = (rng.random(n_reads) * (maxstart + 1)).astype(np.int64)
    strand = rng.random(n_reads) < 0.5
    read_id = np.repeat(np.arange(n_reads), lens)
    within = np.arange(total, dtype=np.int64) - offsets[:-1].astype(np.int64)[read_id]
    L = lens[read_id]
    fwd_idx = gstart[gi][read_id] + st[read_id] + within
    rev_idx = gstart[gi][read_id] + st[read_id] + (L - 1 - within)
    idx = np.where(strand[read_id], rev_idx, fwd_idx)
    idx = np.minimum(idx, len(gcat) - 1)
    b = gcat[idx]
    b = np.where(strand[read_id], _COMP[b], b)
    rnd = BASES[rng.integers(0, 4, size=total)]
    bases = np.where(from_db[read_id], b, rnd).astype(np.uint8)
    if sub_rate > 0:
        m = rng.random(total) < sub_rate
        bases[m] = BASES[rng.integers(0, 4, size=int(m.sum()))]
    if n_rate > 0:
        m = rng.random(total) < n_rate
        bases[m] = ord("N")
    src = np.where(from_db, gi, -1).astype(np.int32)
    return bases, offsets, src


def fastq_bytes(bases, offsets, src=None, qual=b"I", prefix="r"):
    """FASTQ text with `@r<i> <src>` headers and constant quality."""
    out = []
    n = len(offsets) - 1
    bb = bases.tobytes()
    for i in range(n):
        a, b = int(offsets[i]), int(offsets[i + 1])
        out.append(b"@%s%d %d\n" % (prefix.encode(), i, -1 if src is None else int(src[i])))
        out.append(bb[a:b])
        out.append(b"\n+\n")
        out.append(qual * (b - a))
        out.append(b"\n")
    return b"".join(out)


def parent_array_fanout(levels=4, fanout=10):
    """Value-index tree of the bench DBs: index 0 = root, then level by level; children of node p at level l are
    contiguous.  Returns (parent int32[V], first index of every level)."""
    parents = [-1]
    level_start = [0]
    prev = [0]
    for _ in range(levels):
        level_start.append(len(parents))
        cur = []
        for p in prev:
            for _ in range(fanout):
                cur.append(len(parents))
                parents.append(p)
        prev = cur
    return np.array(parents, dtype=np.int32), level_start
