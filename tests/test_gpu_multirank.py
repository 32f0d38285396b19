"""One process per GPU (the torchrun shape of bench.py --gpus N): every rank builds its own replica of the database, matches its
share of the reads, and gs_match_finish_comm merges the per-GPU state inside the library -- counters / max-contigs over NCCL,
unique-k-mer bitsets OR-merged over NVLink peer mappings (and, forced with GS_MERGE_PATH=nccl, over an ncclSend/ncclRecv slice
exchange).  The merged result on EVERY rank must equal the oracle's single pass over all reads
(KMerUniqueCounterBits.getUniqueKmerCounts, C/store/KMerUniqueCounterBits.java:146-163; FastqKMerMatcher.runMatcher :199-234).
Needs two GPUs; a one-GPU box skips it (the host-side logic is covered on the CPU by test_dist_cpu.py)."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = 31
CONFIGS = (dict(), dict(max_kmer_res_counts=3), dict(layout=1), dict(count_unique_kmers=0))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, merge_path, out):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        if merge_path:
            os.environ["GS_MERGE_PATH"] = merge_path
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", rank=rank, world_size=world)   # only carries the 128-byte communicator id
        import gs_oracle
        import util
        from genestrip_b200 import capi, synth
        from genestrip_b200.dist import open_comm
        ctx = capi.Context([rank])
        comm = open_comm(dist, capi, ctx)
        nodes, names, genomes = util.small_project(genome_len=30000, seed=41)
        odb, gdb = util.build_pair(gs_oracle, capi, ctx, K, nodes, names, genomes)   # same store on every rank -> same slot ids
        bases, offsets, src = synth.sample_reads([g for _, g in genomes], 6000, 150, seed=8, frac_db=0.7, sub_rate=0.01, n_rate=0.002)
        n = len(offsets) - 1
        lo, hi = rank * n // world, (rank + 1) * n // world        # this rank's share; read ordinals stay global
        report = []
        for ci, cfg in enumerate(CONFIGS):
            sess = capi.MatchSession(gdb, capi.default_match_cfg(**cfg))
            if ci % 2 == 0:   # with and without the set-up done ahead of the run (gs_match_prepare_merge)
                sess.prepare_merge(comm)
            res = []
            for b0 in range(lo, hi, 700):
                b1 = min(hi, b0 + 700)
                t = sess.submit(bases, np.ascontiguousarray(offsets[b0:b1 + 1]), b0)
                res.append(sess.collect(t)[0])
            counts, top = sess.finish(comm)
            stats = sess.merge_stats()
            sess.close()
            orun = odb.match_files(util.oracle_cfg(gs_oracle, K, **cfg), [synth.fastq_bytes(bases, offsets, src)])
            all_res = np.concatenate(res)
            # per-read results of the own share, merged per-taxon results of ALL reads -- on every rank
            o = orun.reads
            np.testing.assert_array_equal(all_res["class_vidx"], o["class_vidx"][lo:hi])
            fake = np.zeros(n, dtype=capi.READ_RESULT_DTYPE)
            fake["class_vidx"] = o["class_vidx"]; fake["tax_err"] = np.where(o["tax_err"] < 0, 0xFFFFFFFF, o["tax_err"]).astype(np.uint32)
            fake["flags"] = np.where(o["accepted"] != 0, capi.GS_READ_ACCEPTED, 0); fake["read_kmers"] = o["read_kmers"]
            util.assert_match_parity(capi, orun, fake, counts, top, check_unique=bool(cfg.get("count_unique_kmers", 1)))
            first = np.array([int(x) for x in orun.stats["maxlen"]])
            assert (counts["max_contig_len"] == first).all()
            report.append((dict(cfg), stats))
        comm.close()
        gdb.close()
        odb.free()
        ctx.close()
        dist.destroy_process_group()
        out.put((rank, "ok", report))
    except Exception:
        out.put((rank, "error", traceback.format_exc()))


@pytest.mark.parametrize("merge_path", ["", "nccl"])
def test_one_process_per_gpu_merge_matches_single_pass(native, oracle, merge_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2); the merge's host logic is covered by test_dist_cpu.py")
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, merge_path, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, status, info in sorted(got):
        assert status == "ok", "rank %d:\n%s" % (rank, info)
        for cfg, (total_ms, bitset_ms, nbytes, path) in info:
            print("rank %d cfg %s: merge %.3f ms (bitset %.3f ms, %d bytes from peers, path %d)" % (rank, cfg, total_ms, bitset_ms, nbytes, path))
            if cfg.get("count_unique_kmers", 1):
                assert path == (2 if merge_path == "nccl" else 1)


def test_table_layout_is_a_function_of_the_key_set(native, oracle, gpu_ctx):
    """Two independent builds of the same store give the same slot id to every k-mer (what the cross-rank bitset OR relies on):
    the table-layout storage positions of a label dump agree position by position."""
    import torch
    import util
    from genestrip_b200 import synth
    nodes, names, genomes = util.small_project(genome_len=40000, seed=3)
    odb = oracle.OracleDb.build(K, nodes, names, genomes)
    dbs = [util.upload(oracle, native, gpu_ctx, odb) for _ in range(2)]
    try:
        bases, offsets, _ = synth.sample_reads([g for _, g in genomes], 3000, 150, seed=5, frac_db=0.9, sub_rate=0.002, n_rate=0.0)
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        koff = np.zeros(len(offsets), dtype=np.uint64)
        koff[1:] = np.cumsum(np.maximum(lens - K + 1, 0))
        dev = torch.device("cuda:0")
        pad = np.zeros(64, dtype=np.uint8)
        d_b = torch.from_numpy(np.concatenate([bases, pad])).to(dev)
        d_o = torch.from_numpy(offsets.astype(np.int64)).to(dev)
        d_k = torch.from_numpy(koff.astype(np.int64)).to(dev)
        pos = []
        for db in dbs:
            sess = native.MatchSession(db, native.default_match_cfg(count_unique_kmers=0))
            d_l = torch.empty(int(koff[-1]), dtype=torch.int32, device=dev)
            d_p = torch.empty(int(koff[-1]), dtype=torch.int64, device=dev)
            sess.dump_labels(d_b.data_ptr(), d_o.data_ptr(), len(offsets) - 1, d_k.data_ptr(), d_l.data_ptr(), d_p.data_ptr())
            sess.close()
            pos.append(d_p.cpu().numpy())
            assert (d_l.cpu().numpy() >= 0).sum() > 100000
        np.testing.assert_array_equal(pos[0], pos[1])
        hit = pos[0][pos[0] >= 0]
        assert len(np.unique(hit)) > 50000   # many distinct slots took part
    finally:
        for db in dbs:
            db.close()
        odb.free()
