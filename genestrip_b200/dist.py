"""One process per GPU: joining the ranks of a torch.distributed job to the library's own communicator.

Reads shard over ranks with the database replicated, so nothing is exchanged while matching.  The end-of-run merge itself
(counters by ncclAllReduce(sum), max-contigs by ncclAllReduce(max), unique-k-mer bitsets OR-merged slice-wise over NVLink and
counted per taxon, include/genestrip_b200.h "several GPUs") lives behind the C ABI: `MatchSession.finish(comm)` =
gs_match_finish_comm.  What is left here is the one thing the library cannot do by itself: hand rank 0's 128-byte
communicator id to the other processes.  bench.py and the tests use the torch.distributed process group for that (gloo or
NCCL); the Java host would use its own channel.

`merge_match_state_reference` is a plain torch.distributed statement of the same merge.  It is NOT on the product path: the
world_size-2 gloo tests (tests/test_dist_cpu.py) pin the merge semantics with it on the CPU, and the 2-GPU test compares the
library's merge against it.
"""
import torch


def broadcast_unique_id(dist, make_id):
    """Rank 0 calls make_id() (-> 128 bytes); every rank returns the same bytes."""
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])


def open_comm(dist, capi, ctx):
    """gs_comm for this rank of the process group `dist` (collective)."""
    uid = broadcast_unique_id(dist, capi.Comm.unique_id)
    return capi.Comm(ctx, uid, dist.get_world_size(), dist.get_rank())


def slice_bounds(n_words, world, rank):
    """Words [lo, hi) of the bitset that `rank` merges; slices start at multiples of 64 words (merge_slice in gs_capi.cu)."""
    per = (n_words + world - 1) // world
    per = (per + 63) & ~63
    lo = min(n_words, rank * per)
    return per, lo, min(n_words, lo + per)


def merge_match_state_reference(dist, counters, maxcontig, bitset, n_values, popcount_slice):
    """In place on counters / maxcontig; returns unique[V] (int64, summed over ranks) or None without a bitset.

    popcount_slice(merged_words, word_lo, word_hi) -> int64[V]: per-taxon popcount of the OR-merged words [word_lo, word_hi).
    NCCL has no bitwise OR and `max` on bytes is OR only for one flag per byte, hence: exchange of 1/N slices, local OR,
    per-taxon popcount of the own slice, all_reduce(SUM) of the [V] counts.
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
