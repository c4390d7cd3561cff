"""-m gpu: dkb_finalise_from and the allreduce/scan overlap helper (world size 1: the sum is a no-op)."""
import numpy as np
import pytest

from denovo_kmer_b200 import dist, synth
from denovo_kmer_b200.api import DEFAULT_THRESHOLDS

pytestmark = pytest.mark.gpu


def _submit_all(dkb, kc, trio, k):
    for smp in range(3):
        seq, qual, off = trio.reads[smp]
        kc.submit(dkb.pack_reads(seq, qual, off, 20), smp)


def test_count_pipeline_matches_plain_finalise(dkb):
    k = 31
    trio = synth.make_trio_host(100_000, 10, 20, k, seed=3)
    entries = dkb.variant_kmers(trio.variant_tuples(), k)
    with dkb.KmerCounter(k) as kc:
        kc.build_table(entries)
        _submit_all(dkb, kc, trio, k)
        plain = kc.finalise(DEFAULT_THRESHOLDS)
        counts = kc.entry_counts().copy()
        assert counts.sum() > 0
        pipe = dist.CountPipeline(kc, DEFAULT_THRESHOLDS)
        for _ in range(3):  # three batches in flight one after the other; results of the last one
            kc.reset_counts()
            _submit_all(dkb, kc, trio, k)
            pipe.push()
        pipe.flush()
        res = kc.results()
        for a, b in zip(plain, res):
            assert np.array_equal(a, b)
        assert np.array_equal(kc.reduced_counts(), counts)
        assert np.array_equal(kc.entry_counts(), counts)
        assert kc.comm_info()[:2] == (0, 1)
        kc.counts_allreduce()  # no communicator: a no-op
        assert np.array_equal(kc.entry_counts(), counts)


def test_reduce_snapshots_survive_the_next_batch(dkb):
    """The snapshot of batch i is finalised AFTER batch i + 1 has been scanned into the live
    counters: results must be batch i's, not a mix."""
    k = 31
    t1 = synth.make_trio_host(100_000, 10, 20, k, seed=3)
    entries = dkb.variant_kmers(t1.variant_tuples(), k)
    with dkb.KmerCounter(k) as kc:
        kc.build_table(entries)
        _submit_all(dkb, kc, t1, k)
        want = kc.finalise(DEFAULT_THRESHOLDS)
        c1 = kc.entry_counts().copy()
        kc.reset_counts()
        _submit_all(dkb, kc, t1, k)
        kc.reduce_push(DEFAULT_THRESHOLDS)           # batch 1
        kc.reset_counts()
        for smp in (0, 0, 1):                         # batch 2: different work into the live counters
            seq, qual, off = t1.reads[smp]
            kc.submit(dkb.pack_reads(seq, qual, off, 20), smp)
        kc.reduce_push(DEFAULT_THRESHOLDS)           # finalises batch 1 behind batch 2's scans
        got = kc.results()
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
        assert np.array_equal(kc.reduced_counts(), c1)
        kc.reduce_flush(DEFAULT_THRESHOLDS)          # batch 2
        c2 = kc.reduced_counts()
        assert np.array_equal(c2[0], 2 * c1[0]) and np.array_equal(c2[1], c1[1]) and c2[2].sum() == 0


def test_finalise_from_other_counts(dkb):
    """Kernel 3 on a caller-owned copy: doubling every count doubles the hits."""
    import torch
    k = 21
    trio = synth.make_trio_host(60_000, 8, 20, k, seed=4)
    entries = dkb.variant_kmers(trio.variant_tuples(), k)
    with dkb.KmerCounter(k) as kc:
        kc.build_table(entries)
        _submit_all(dkb, kc, trio, k)
        hits, distinct, n_kmers, _ = kc.finalise(DEFAULT_THRESHOLDS)
        t = dist.counts_tensor(kc)
        kc.sync()
        twice = (t * 2).contiguous()
        torch.cuda.synchronize()
        kc.finalise_launch(DEFAULT_THRESHOLDS, counts_ptr=twice.data_ptr())
        hits2, distinct2, n_kmers2, _ = kc.results()
        assert np.array_equal(hits2, hits * 2)
        assert np.array_equal(distinct2, distinct) and np.array_equal(n_kmers2, n_kmers)
