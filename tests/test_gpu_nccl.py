"""-m gpu, needs 2 GPUs (skipped on one): the real multi-GPU path — two processes, one GPU
each, their own shards of every sample's reads, the NCCL communicator and the overlapped
reduction inside libdkb.so (dkb_comm_init, dkb_reduce_push / _flush, dkb_counts_allreduce).
Four pipelined steps with different reads per step; every step's summed counters, per-variant
statistics and calls must equal the oracle's over the union of both ranks' reads.  The NCCL
id travels through a file: no torch.distributed anywhere."""
import os
import sys
import time

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

K, THR, STEPS = 31, (3, 2, 0, 1), 4


def _worker(rank, world, tmp):
    sys.path.insert(0, ROOT)
    import denovo_kmer_b200 as dkb
    from denovo_kmer_b200 import dist, synth
    import oracle
    idf = os.path.join(tmp, "nccl_id")
    if rank == 0:
        with open(idf + ".tmp", "wb") as f:
            f.write(dkb.comm_unique_id())
        os.rename(idf + ".tmp", idf)
    while not os.path.exists(idf):
        time.sleep(0.01)
    uid = open(idf, "rb").read()
    trios = [synth.make_trio_host(120_000, 12, 24, K, seed=200 + s) for s in range(STEPS)]
    # one table for all steps: step 0's candidates; later steps reuse its genome with new reads
    genome, variants = trios[0].genome, trios[0].variants
    entries = dkb.variant_kmers(trios[0].variant_tuples(), K)
    ks = oracle.KmerSet(entries.keys, entries.variant, entries.allele)
    child_alt = synth.apply_variants(genome, variants)
    reads = [[synth.sample_reads([genome, child_alt] if smp == 0 else [genome, genome], 9000, 150,
                                 seed=1000 + 10 * s + smp) for smp in range(3)] for s in range(STEPS)]
    out = {}
    with dkb.KmerCounter(K, device=rank) as kc:
        kc.comm_init(uid, rank, world)
        assert kc.comm_info()[:2] == (rank, world)
        kc.build_table(entries)
        got = []
        for s in range(STEPS):
            kc.reset_counts()
            for smp in range(3):
                seq, qual, off = reads[s][smp]
                lo, hi = dist.shard_range(len(off) - 1, rank, world)   # this rank's reads
                a, b = int(off[lo]), int(off[hi])
                kc.submit(dkb.pack_reads(seq[a:b], qual[a:b], off[lo:hi + 1] - off[lo], 20), smp)
            kc.reduce_push(THR)
            if s > 0:  # the push finalised step s - 1 behind step s's scans
                got.append((kc.reduced_counts().copy(),) + tuple(x.copy() for x in kc.results()))
        kc.reduce_flush(THR)
        got.append((kc.reduced_counts().copy(),) + tuple(x.copy() for x in kc.results()))
        # serial form on the last step: in-place allreduce of the live counters, then finalise
        kc.counts_allreduce()
        serial = (kc.entry_counts().copy(),) + tuple(x.copy() for x in kc.finalise(THR))
        kc.comm_destroy()
    for s in range(STEPS):
        want = np.zeros((3, len(entries)), dtype=np.uint64)
        for smp in range(3):
            seq, qual, off = reads[s][smp]
            ks.count_reads(seq, qual, off, K, 20, counts=want[smp])   # the union of both shards
        h, d, nk = ks.variant_stats(want, entries.n_variants)
        calls = oracle.calls(h, d, THR)
        c, gh, gd, gnk, gcalls = got[s]
        assert np.array_equal(c.astype(np.uint64), want), f"rank {rank} step {s}: counters"
        assert np.array_equal(gh.astype(np.uint64), h) and np.array_equal(gd.astype(np.uint64), d)
        assert np.array_equal(gnk, nk) and np.array_equal(gcalls, calls)
        assert want.sum() > 0
        if s == STEPS - 1:
            assert np.array_equal(serial[0].astype(np.uint64), want) and np.array_equal(serial[4], calls)
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("ok")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_pipelined_counts(tmp_path):
    mp.spawn(_worker, args=(2, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
