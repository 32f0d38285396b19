"""One process driving two GPUs through ONE context (gs_ctx_create with two ordinals): the database is replicated, batches are
dealt round-robin, gs_match_finish merges counters / max-contigs / unique-k-mer bits across the devices.  Needs two GPUs;
the single-GPU tiers skip it (the torchrun path over NCCL is covered by test_dist_cpu.py and bench.py --gpus N)."""
import numpy as np
import pytest

from genestrip_b200 import synth

import util

pytestmark = pytest.mark.gpu

K = 31


@pytest.fixture(scope="module")
def ctx2(native):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    ctx = native.Context([0, 1])
    yield ctx
    ctx.close()


def test_two_devices_one_context(oracle, native, ctx2):
    nodes, names, genomes = util.small_project(genome_len=30000, seed=41)
    odb, gdb = util.build_pair(oracle, native, ctx2, K, nodes, names, genomes)
    try:
        bases, offsets, src = synth.sample_reads([g for _, g in genomes], 6000, 150, seed=8, frac_db=0.7, sub_rate=0.01, n_rate=0.002)
        fq = synth.fastq_bytes(bases, offsets, src)
        for cfg in (dict(), dict(max_kmer_res_counts=3), dict(layout=1)):
            orun = odb.match_files(util.oracle_cfg(oracle, K, **cfg), [fq])
            res, ev, counts, top, _, launches = util.gpu_match(native, gdb, bases, offsets, batch=500, **cfg)   # 12 batches over 2 devices
            util.assert_match_parity(native, orun, res, counts, top)
        # FASTQ text chunks alternate between the devices as well
        recs = fq.split(b"\n@")
        recs = [recs[0] + b"\n"] + [b"@" + r + (b"\n" if not r.endswith(b"\n") else b"") for r in recs[1:]]
        sess = native.MatchSession(gdb, native.default_match_cfg())
        try:
            out, pend, ordinal = [], [], 0
            for a in range(0, len(recs), 750):
                chunk = np.frombuffer(b"".join(recs[a:a + 750]), dtype=np.uint8)
                t, info = sess.submit_fastq(chunk, ordinal)
                assert t and info.status == 0
                ordinal += info.n_reads
                pend.append(t)
                if len(pend) == 2 * native.GS_MAX_INFLIGHT:
                    out.append(sess.collect_fastq(pend.pop(0))[0].copy())
            while pend:
                out.append(sess.collect_fastq(pend.pop(0))[0].copy())
            counts, top = sess.finish()
        finally:
            sess.close()
        orun = odb.match_files(util.oracle_cfg(oracle, K), [fq])
        util.assert_match_parity(native, orun, np.concatenate(out), counts, top)
    finally:
        gdb.close()
        odb.free()
