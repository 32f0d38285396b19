"""world_size-2 (and 3) gloo tests of the multi-rank host logic (no GPU needed): the communicator id reaches every rank, the
slice geometry covers the bitset, and the merge semantics the library implements on the GPUs (gs_match_finish_comm) are pinned
by their plain torch.distributed statement."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_state(rank, V, n_pos, seed=7):
    rng = np.random.default_rng(seed + rank)
    counters = rng.integers(0, 1000, size=(7, V)).astype(np.int64)
    lens = rng.integers(0, 120, size=V).astype(np.int64)
    ordinals = rng.integers(0, 1 << 30, size=V).astype(np.int64)
    maxcontig = np.where(lens > 0, (lens << 40) | ((1 << 40) - 1 - ordinals), 0)
    bits = (rng.random(n_pos) < 0.3)
    return counters, maxcontig, bits


def _pack(bits):
    n_words = (len(bits) + 63) // 64
    padded = np.zeros(n_words * 64, dtype=bool)
    padded[: len(bits)] = bits
    return np.packbits(padded.reshape(-1, 8)[:, ::-1], axis=None).view("<u8").astype(np.uint64).view(np.int64) if False else \
        (padded.reshape(n_words, 64).astype(np.uint64) << np.arange(64, dtype=np.uint64)).sum(axis=1).astype(np.uint64).view(np.int64)


def _worker(rank, world, port, V, n_pos, vals, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genestrip_b200.dist import merge_match_state_reference as merge_match_state, broadcast_unique_id
    uid = broadcast_unique_id(dist, lambda: bytes(range(128)))
    assert uid == bytes(range(128))
    counters, maxcontig, bits = _rank_state(rank, V, n_pos)
    t_c, t_m, t_b = torch.from_numpy(counters.copy()), torch.from_numpy(maxcontig.copy()), torch.from_numpy(_pack(bits).copy())

    def popcount_slice(words, lo, hi):  # numpy stand-in of gs_unique_popcount_kernel for the CPU test
        w = words.numpy().view(np.uint64)
        uniq = np.zeros(V, dtype=np.int64)
        for i, word in enumerate(w):
            for b in range(64):
                if (int(word) >> b) & 1:
                    pos = (lo + i) * 64 + b
                    if pos < n_pos:
                        uniq[vals[pos]] += 1
        return torch.from_numpy(uniq)

    unique = merge_match_state(dist, t_c, t_m, t_b, V, popcount_slice)
    if rank == 0:
        out.put((t_c.numpy().copy(), t_m.numpy().copy(), unique.numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_pos", [(2, 1000), (3, 777), (2, 64)])
def test_merge_match_state_gloo(world, n_pos):
    V = 9
    vals = np.random.default_rng(3).integers(0, V, size=n_pos)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, V, n_pos, vals, out)) for r in range(world)]
    for p in procs:
        p.start()
    got_c, got_m, got_u = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    states = [_rank_state(r, V, n_pos) for r in range(world)]
    np.testing.assert_array_equal(got_c, sum(s[0] for s in states))
    np.testing.assert_array_equal(got_m, np.maximum.reduce([s[1] for s in states]))
    merged = np.logical_or.reduce([s[2] for s in states])
    exp_u = np.bincount(vals[merged], minlength=V)
    np.testing.assert_array_equal(got_u, exp_u)
    # ties on the contig length resolve to the lowest read ordinal
    lens = got_m >> 40
    for v in range(V):
        cands = [((1 << 40) - 1 - (s[1][v] & ((1 << 40) - 1))) for s in states if (s[1][v] >> 40) == lens[v] and s[1][v]]
        if cands:
            assert (1 << 40) - 1 - (got_m[v] & ((1 << 40) - 1)) == min(cands)


def test_slice_bounds_cover_everything():
    from genestrip_b200.dist import slice_bounds
    for n_words in (0, 1, 7, 64, 1000, 8191, 1 << 20):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                per, lo, hi = slice_bounds(n_words, world, r)
                covered.extend(range(lo, hi))
                assert hi - lo <= per and (lo % 64 == 0 or lo == hi)
            assert covered == list(range(n_words))
