"""-m gpu: BASELINE.json's configurations AT THE SIZE bench.py runs them, bit-exact against
the CPU oracle on every per-entry counter, per-variant statistic and call byte.

The streams are generated in HBM by the bench's own generator (same seeds as bench.py rank
0), scanned by the CUDA path (auto-tuned, all three samples in one launch), then copied to
the host and counted by oracle.count_stream — the plain per-position rolling-window walk of
the packed stream, itself pinned to the per-read oracle by tests/test_oracle.py.
  configs[1]  64 Mb, 30x, 10 000 candidates, k=31 — must tune to (15, 16, 2, L2) + 37 120-word pre-filter
  configs[2]  one GPU's shard: 128 Mb of 30x reads, 4 000 local of a 100 000-candidate table
  configs[3]  64 Mb, 100x, 50 000 candidates incl. indels (19.2 Gbases)
  configs[4]  k = 15 / 21 / 25 / 31 on 64 Mb, 30x, 10 000 candidates, base-quality masking on
"""
import argparse

import numpy as np
import pytest
import torch

import bench
from denovo_kmer_b200 import synth

pytestmark = pytest.mark.gpu

THR = (3, 2, 0, 1)


def _args(**kw):
    a = bench.parse_args([])
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def _run(dkb, orc, a, lowq_frac=None):
    dev = torch.device("cuda:0")
    genome, variants, tuples = bench.make_variants(a, synth)
    entries = dkb.variant_kmers(tuples, a.k)
    if lowq_frac is None:
        streams = bench.make_streams(a, synth, genome, variants, dev)
    else:  # heavier base-quality masking than the bench default
        lut = np.zeros(256, dtype=np.uint8)
        for i, ch in enumerate(b"ACGT"):
            lut[ch] = i
        g = torch.from_numpy(lut[genome]).to(dev)
        ca = torch.from_numpy(lut[synth.apply_variants(genome, variants)]).to(dev)
        ma = torch.from_numpy(lut[synth.apply_variants(genome, [v for v in variants if v.inherited])]).to(dev)
        n_reads = int(len(genome) * a.depth / bench.READ_LEN) // 128 * 128
        streams = [synth.make_sample_device(h, n_reads, bench.READ_LEN, 1000 + s, dev, lowq_frac=lowq_frac,
                                            n_rate=0.002) for s, h in enumerate([[g, ca], [g, ma], [g, g]])]
        del g, ca, ma
    torch.cuda.synchronize()
    with dkb.KmerCounter(a.k) as kc:
        kc.build_table(entries)
        per = ((1 << 32) - 4096) // 151 // 128 * 128 * 151   # one launch takes < 2^32 positions
        if all(n_pos <= per for (_, _, n_pos, _) in streams):
            kc.submit_device_multi([(b2.data_ptr(), m1.data_ptr(), n_pos, s)
                                    for s, (b2, m1, n_pos, _) in enumerate(streams)])
        else:  # 100x: 6.4 G positions per sample, cut after a multiple of 128 reads
            for s, (b2, m1, n_pos, _) in enumerate(streams):
                for p0 in range(0, n_pos, per):
                    kc.submit_device(b2.data_ptr() + p0 // 4, m1.data_ptr() + p0 // 8, min(per, n_pos - p0), s)
        got = kc.entry_counts().copy()
        hits, dist, nk, calls = kc.finalise(THR)
        tun, st = kc.tuning(), kc.stats()
    ks = orc.KmerSet(entries.keys, entries.variant, entries.allele)
    want = np.zeros((3, len(entries)), dtype=np.uint64)
    for s, (b2, m1, n_pos, _) in enumerate(streams):
        ks.count_stream(b2.cpu().numpy(), m1.cpu().numpy(), n_pos, a.k, counts=want[s])
    del streams
    torch.cuda.empty_cache()
    bad = int((got.astype(np.uint64) != want).sum())
    assert bad == 0, f"{bad} of {want.size} counters differ from the oracle (tuning {tun})"
    o_hits, o_dist, o_nk = ks.variant_stats(want, entries.n_variants)
    assert np.array_equal(hits.astype(np.uint64), o_hits) and np.array_equal(dist.astype(np.uint64), o_dist)
    assert np.array_equal(nk, o_nk)
    assert np.array_equal(calls, orc.calls(o_hits, o_dist, THR))
    assert want.sum() > 1_000_000
    return tun, st, calls, variants


def test_config1_64mb_30x_10k_full_size(dkb, orc):
    tun, st, calls, variants = _run(dkb, orc, _args())
    assert tun == (15, 16, 2, 2) and st["prefilter_words"] == 37120, (tun, st["prefilter_words"])
    inherited = np.array([v.inherited for v in variants])
    assert (calls[~inherited] & 1).mean() > 0.9 and (calls[inherited] & 1).sum() == 0


def test_config2_wgs_shard_100k_table_full_size(dkb, orc):
    tun, st, _, _ = _run(dkb, orc, _args(genome_mb=128.0, variants=4000, table_variants=100000))
    assert tun[1] == 16 and tun[3] == 2 and st["n_entries"] == 6_200_000


def test_config3_100x_50k_indels_full_size(dkb, orc):
    _run(dkb, orc, _args(depth=100.0, variants=50000, indel_frac=0.5))


@pytest.mark.parametrize("k", [15, 21, 25, 31])
def test_config4_k_sweep_bq_masking_full_size(dkb, orc, k):
    _run(dkb, orc, _args(k=k), lowq_frac=0.08)
