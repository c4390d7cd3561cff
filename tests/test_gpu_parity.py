"""-m gpu: the CUDA path through the C ABI against the CPU oracle, bit-exact."""
import numpy as np
import pytest

from denovo_kmer_b200 import synth
from helpers import gpu_counts, oracle_counts

pytestmark = pytest.mark.gpu


def _check(dkb, orc, trio, k, min_bq=20, drop_shared=True, **kw):
    entries = dkb.variant_kmers(trio.variant_tuples(), k, drop_shared=drop_shared)
    ks, want = oracle_counts(orc, entries, trio, k, min_bq)
    got, stats, tun = gpu_counts(dkb, entries, trio, k, min_bq, **kw)
    assert want.max() < 2 ** 32
    bad = np.nonzero(got.astype(np.uint64) != want)
    assert len(bad[0]) == 0, (
        f"tuning={tun} mismatches={len(bad[0])} first={[(int(s), int(e), int(got[s, e]), int(want[s, e])) for s, e in zip(*bad)][:5]}")
    assert want.sum() > 0
    return entries, ks, want, got, stats


@pytest.mark.parametrize("tuning", [None, (15, 1, 1), (15, 1, 2), (13, 1, 2), (15, 2, 1), (15, 2, 2),
                                    (14, 4, 1), (14, 4, 2), (12, 1, 1), (9, 2, 1), (15, 8, 2), (13, 8, 1), (15, 16, 2),
                                    (9, 16, 1), (15, 16, 2, 2), (14, 4, 1, 2), (15, 2, 2, 2), (13, 8, 2, 2)])
@pytest.mark.parametrize("hints", [True, False])
def test_snv_trio_k31(dkb, orc, tuning, hints):
    trio = synth.make_trio_host(200_000, 12, 40, 31, seed=5)
    _check(dkb, orc, trio, 31, tuning=tuning, hints=hints)


@pytest.mark.parametrize("gate", ["0", "1"])
@pytest.mark.parametrize("pre_words", [0, 1024, 37120])
@pytest.mark.parametrize("tuning", [(15, 16, 2, 2), (15, 16, 1, 2), (14, 8, 2, 2), (13, 8, 1, 2), (9, 16, 2, 2)])
def test_gated_lookups(dkb, orc, tuning, pre_words, gate, monkeypatch):
    """Strides 8 and 16 of the L2 filter modes with the lookups gated by the flag stream (a seed
    that holds an N, a low-quality base or a separator is not looked up) forced on and off: the
    counters are the oracle's either way; heavy masking so that most seeds are affected."""
    monkeypatch.setenv("DKB_PREFILTER_WORDS", str(pre_words))
    monkeypatch.setenv("DKB_GATE", gate)
    trio = synth.make_trio_host(150_000, 14, 30, 31, seed=77, n_rate=0.01, lowq_frac=0.08, ragged=True)
    _check(dkb, orc, trio, 31, tuning=tuning)
    trio = synth.make_trio_host(150_000, 14, 30, 31, seed=78, n_rate=0.0, lowq_frac=0.0)
    _check(dkb, orc, trio, 31, tuning=tuning, hints=False)


@pytest.mark.parametrize("pre_words", [0, 64, 1024, 32768, 35584])
@pytest.mark.parametrize("tuning", [(15, 16, 2, 2), (15, 16, 1, 2), (14, 8, 2, 2), (14, 4, 2, 2), (12, 2, 1, 2)])
def test_l2_filter_behind_prefilter(dkb, orc, tuning, pre_words, monkeypatch):
    """L2 filter mode with the shared-memory pre-filter forced to several sizes (64 words:
    saturated, passes nearly everything; 35584: the size the tuner uses); 0 = pure L2 mode."""
    monkeypatch.setenv("DKB_PREFILTER_WORDS", str(pre_words))
    trio = synth.make_trio_host(200_000, 12, 40, 31, seed=5)
    _check(dkb, orc, trio, 31, tuning=tuning)
    _check(dkb, orc, trio, 31, tuning=tuning, hints=False)


def test_prefilter_chosen_by_tuner(dkb):
    """A table of configs[1]'s size makes the tuner pick the L2 filter at stride 16 (behind the
    pre-filter); 250 candidates keep the shared-memory filter."""
    for n_var, want_mode in [(10_000, 2), (250, 1)]:
        genome = synth.make_genome(max(4_000_000, n_var * 400), 1)
        variants = synth.plant_variants(genome, n_var, 31, 2)
        entries = dkb.variant_kmers(synth.Trio(31, genome, variants).variant_tuples(), 31)
        with dkb.KmerCounter(31) as kc:
            kc.build_table(entries)
            tun = kc.tuning()
        assert tun[1] == 16 and tun[3] == want_mode, tun


@pytest.mark.parametrize("k", [15, 21, 25, 31, 8, 16, 30])
def test_k_sweep(dkb, orc, k):
    trio = synth.make_trio_host(100_000, 10, 20, k, seed=7 + k)
    _check(dkb, orc, trio, k)
    _check(dkb, orc, trio, k, hints=False, tuning=(0, 2 if k > 8 else 1, 0))


def test_indels_and_shared(dkb, orc):
    trio = synth.make_trio_host(150_000, 15, 40, 31, seed=11, indel_frac=0.6)
    _check(dkb, orc, trio, 31)
    _check(dkb, orc, trio, 31, drop_shared=False)
    _check(dkb, orc, trio, 31, drop_shared=False, hints=False, tuning=(15, 2, 2))


def test_ragged_reads_and_batches(dkb, orc):
    trio = synth.make_trio_host(100_000, 15, 20, 21, seed=13, ragged=True, n_rate=0.01,
                                lowq_frac=0.1)
    _check(dkb, orc, trio, 21, batches=7)


def test_no_quality_no_masking(dkb, orc):
    trio = synth.make_trio_host(100_000, 10, 20, 31, seed=17)
    _check(dkb, orc, trio, 31, min_bq=0)
