"""-m gpu: BASELINE.json's configs as parity cases against the oracle, bit-exact on every
per-entry counter, per-variant statistic and call byte.
  configs[0]  1 Mb region, 30x, 150 bp, 100 candidate DNMs, k=31          — at full size
  configs[3]  100x high depth, 50k candidates incl. indels per 64 Mb      — same depth and
              candidate density on a 2 Mb region (1560 candidates)
  configs[4]  k sweep 15/21/25/31 with base-quality masking, 10k per 64 Mb — same density
              on a 1 Mb region (156 candidates)
configs[1] / [2] are bench.py's workloads; tests/test_gpu_fullsize.py checks their shape
against analytic truth."""
import numpy as np
import pytest

from denovo_kmer_b200 import synth
from helpers import gpu_counts, oracle_counts

pytestmark = pytest.mark.gpu


def _full_check(dkb, orc, trio, k, min_bq=20, thresholds=(3, 2, 0, 1)):
    entries = dkb.variant_kmers(trio.variant_tuples(), k)
    ks, want = oracle_counts(orc, entries, trio, k, min_bq)
    with dkb.KmerCounter(k) as kc:
        kc.build_table(entries)
        for smp in range(3):
            seq, qual, off = trio.reads[smp]
            kc.submit(dkb.pack_reads(seq, qual, off, min_bq), smp)
        got = kc.entry_counts()
        hits, dist, nk, calls = kc.finalise(thresholds)
        tun = kc.tuning()
    assert np.array_equal(got.astype(np.uint64), want), tun
    o_hits, o_dist, o_nk = ks.variant_stats(want, len(trio.variants))
    assert np.array_equal(hits.astype(np.uint64), o_hits)
    assert np.array_equal(dist.astype(np.uint64), o_dist)
    assert np.array_equal(nk, o_nk)
    o_calls = orc.calls(o_hits, o_dist, thresholds)
    assert np.array_equal(calls, o_calls)
    return calls, trio


def test_config0_1mb_30x_100dnm_k31(dkb, orc):
    trio = synth.make_trio_host(1_000_000, 30, 100, 31, seed=1001)
    calls, trio = _full_check(dkb, orc, trio, 31)
    # the planted truth: de novo variants are called, inherited ones are not
    inherited = np.array([v.inherited for v in trio.variants])
    assert (calls[~inherited] & 1).mean() > 0.9
    assert (calls[inherited] & 1).sum() == 0


def test_config3_shape_100x_indels(dkb, orc):
    trio = synth.make_trio_host(2_000_000, 100, 1560, 31, seed=1003, indel_frac=0.5)
    _full_check(dkb, orc, trio, 31)


@pytest.mark.parametrize("k", [15, 21, 25, 31])
def test_config4_shape_k_sweep_bq_masking(dkb, orc, k):
    trio = synth.make_trio_host(1_000_000, 30, 156, k, seed=1004 + k, lowq_frac=0.08,
                                n_rate=0.002)
    _full_check(dkb, orc, trio, k, min_bq=20)
