"""End-of-job merge of the per-rank match state (one process per GPU; the only exchange step of the path).

Reads shard over ranks with the database replicated, so nothing is exchanged while matching.  At the end:
  counters   int64[7][V]  -> all_reduce(SUM)
  maxcontig  int64[V]     -> all_reduce(MAX)   packed (len << 40 | 2^40-1-ordinal) < 2^63, ties -> lowest read ordinal
  bitset     int64[W]     -> NCCL has no bitwise OR and `max` on bytes is OR only for one flag per byte, so: all_to_all of
                             1/N slices, local OR, per-taxon popcount of the own slice, all_reduce(SUM) of the [V] counts
Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch


def slice_bounds(n_words, world, rank):
    per = (n_words + world - 1) // world
    lo = min(n_words, rank * per)
    return per, lo, min(n_words, lo + per)


def merge_match_state(dist, counters, maxcontig, bitset, n_values, popcount_slice):
    """In place on counters / maxcontig; returns unique[V] (int64, summed over ranks) or None without a bitset.

    popcount_slice(merged_words, word_lo, word_hi) -> int64[V]: per-taxon popcount of the OR-merged words [word_lo, word_hi).
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    dist.all_reduce(maxcontig, op=dist.ReduceOp.MAX)
    if bitset is None:
        return None
    n_words = bitset.numel()
    per, lo, hi = slice_bounds(n_words, world, rank)
    padded = torch.zeros(per * world, dtype=torch.int64, device=bitset.device)
    padded[:n_words] = bitset
    recv = torch.empty_like(padded)
    dist.all_to_all_single(recv, padded)
    parts = recv.view(world, per)
    merged = parts[0].clone()
    for r in range(1, world):
        merged |= parts[r]
    unique = popcount_slice(merged[: hi - lo], lo, hi)
    dist.all_reduce(unique, op=dist.ReduceOp.SUM)
    return unique
