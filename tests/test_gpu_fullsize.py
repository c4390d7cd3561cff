"""-m gpu: size-independent properties at a large size (64 Mb-scale is bench.py's job; here
8 Mb x 30x, ~240 M positions per sample, 2000 SNVs) where the oracle would take too long
to be a unit test:
  * analytic truth: for error-free, all-valid reads the count of every entry equals the
    number of reads of the right haplotype that contain its window — computed from the
    read start positions with torch, never touching the kernel under test or the oracle;
  * batch-split invariance and additivity across samples;
  * tuning invariance: every (seed length, stride, hashes) gives identical counters."""
import numpy as np
import pytest
import torch

from denovo_kmer_b200 import synth

pytestmark = pytest.mark.gpu

K = 31
RL = 150


def _setup(dkb, n_var=2000, glen=8_000_000, depth=30):
    dev = torch.device("cuda:0")
    genome = synth.make_genome(glen, 5)
    variants = synth.plant_variants(genome, n_var, K, 6, indel_frac=0.0, inherited_frac=0.0)
    trio = synth.Trio(K, genome, variants)
    entries = dkb.variant_kmers(trio.variant_tuples(), K)
    lut = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    ref = torch.from_numpy(lut[genome]).to(dev)
    alt = torch.from_numpy(lut[synth.apply_variants(genome, variants)]).to(dev)
    n_reads = int(glen * depth / RL) // 128 * 128
    return dev, genome, variants, entries, ref, alt, n_reads


def test_analytic_counts_error_free(dkb):
    dev, genome, variants, entries, ref, alt, n_reads = _setup(dkb)
    b2, m1, n_pos, n_bases, meta = synth.make_sample_device(
        [ref, alt], n_reads, RL, 77, dev, err_rate=0.0, n_rate=0.0, lowq_frac=0.0, rc_frac=0.5,
        return_meta=True)
    with dkb.KmerCounter(K) as kc:
        kc.build_table(entries)
        torch.cuda.synchronize()
        kc.submit_device(b2.data_ptr(), m1.data_ptr(), n_pos, 0)
        got = kc.entry_counts()[0].astype(np.int64)
    # expected: entry with window index u of (variant v, allele a) sits at genome/hap position
    # g = pos_v - (K-1) + u (SNVs keep coordinates); reads of haplotype a starting in
    # [g + K - RL, g] contain it.  Either strand counts (canonical keys).
    hap_of, starts = meta
    pos = torch.tensor([v.pos for v in variants], device=dev)
    ev = torch.from_numpy(entries.variant.astype(np.int64)).to(dev)
    ea = torch.from_numpy(entries.allele.astype(np.int64)).to(dev)
    eu = torch.from_numpy((entries.win_index & 0x7FFF).astype(np.int64)).to(dev)
    g = pos[ev] - (K - 1) + eu
    want = torch.zeros(len(entries), dtype=torch.int64, device=dev)
    for a in (0, 1):
        s = torch.sort(starts[hap_of == a]).values
        hi = torch.searchsorted(s, g, right=True)
        lo = torch.searchsorted(s, g + K - RL, right=False)
        want = torch.where(ea == a, hi - lo, want)
    want = want.cpu().numpy()
    assert np.array_equal(got, want), f"{(got != want).sum()} of {len(got)} entries differ"
    assert want.sum() > 100_000


def test_split_additivity_and_tuning_invariance(dkb):
    dev, genome, variants, entries, ref, alt, n_reads = _setup(dkb, n_var=1500, glen=4_000_000)
    streams = [synth.make_sample_device([ref, alt], n_reads, RL, 80 + s, dev) for s in range(2)]
    torch.cuda.synchronize()
    stride = RL + 1
    results = {}
    for tuning in [None, (15, 1, 1), (15, 1, 2), (15, 2, 1), (15, 2, 2), (14, 4, 1), (14, 4, 2),
                   (14, 4, 3), (12, 2, 4), (9, 1, 1), (15, 8, 2), (15, 16, 1), (12, 16, 2), (15, 16, 2, 2), (14, 4, 2, 2)]:
        with dkb.KmerCounter(K, tuning=tuning) as kc:
            kc.build_table(entries)
            for s, (b2, m1, n_pos, _) in enumerate(streams):
                kc.submit_device(b2.data_ptr(), m1.data_ptr(), n_pos, s)
            whole = kc.entry_counts().copy()
            # same streams cut into 5 batches at multiples of 128 reads, all into sample 2
            kc.reset_counts()
            for (b2, m1, n_pos, _) in streams:
                cuts = [0] + [int(n_reads * f) // 128 * 128 for f in (0.1, 0.37, 0.5, 0.93)] + [n_reads]
                for a, b in zip(cuts[:-1], cuts[1:]):
                    p0, p1 = a * stride, b * stride
                    kc.submit_device(b2.data_ptr() + p0 // 4, m1.data_ptr() + p0 // 8, p1 - p0, 2)
            parts = kc.entry_counts().copy()
        assert np.array_equal(parts[2], whole[0] + whole[1]), tuning
        assert parts[0].sum() == 0 and parts[1].sum() == 0
        results[tuning] = whole
    base = results[None]
    assert base.sum() > 100_000
    for t, r in results.items():
        assert np.array_equal(r, base), t


@pytest.mark.parametrize("tuning", [None, (14, 4, 2, 1), (15, 16, 2, 1), (15, 8, 2, 2)])
def test_maximum_batch_length(dkb, orc, tuning):
    """One batch of the maximum length (2^32 - 4096 positions, 1.07 GB of packed bases):
    random sequence with a real trio's reads packed into its last positions.  Every 32-bit
    position computation near the top of the range must hold: counts equal the oracle's on
    the real reads (a random 31-mer matching the table by chance has probability ~1e-9)."""
    dev = torch.device("cuda:0")
    trio = synth.make_trio_host(50_000, 10, 10, K, seed=91)
    entries = dkb.variant_kmers(trio.variant_tuples(), K)
    seq, qual, off = trio.reads[0]
    tail = dkb.pack_reads(seq, qual, off, 20)
    n_max = (1 << 32) - 4096
    n_head = (n_max - tail.n_positions) // 128 * 128
    n_pos = n_head + tail.n_positions
    bw, mw = dkb.stream_words(n_pos)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    b2 = torch.empty(bw, dtype=torch.int32, device=dev)
    for a in range(0, n_head // 16, 1 << 26):
        e = min(a + (1 << 26), n_head // 16)
        b2[a:e] = torch.randint(-2 ** 31, 2 ** 31 - 1, (e - a,), dtype=torch.int32, device=dev, generator=g)
    m1 = torch.full((mw,), -1, dtype=torch.int32, device=dev)
    tb = torch.from_numpy(tail.bases2.view(np.int32)).to(dev)
    tm = torch.from_numpy(tail.mask1.view(np.int32)).to(dev)
    b2[n_head // 16: n_head // 16 + tb.numel()] = tb[: bw - n_head // 16]
    m1[n_head // 32: n_head // 32 + tm.numel()] = tm[: mw - n_head // 32]
    torch.cuda.synchronize()
    ks = orc.KmerSet(entries.keys, entries.variant, entries.allele)
    want = ks.count_reads(seq, qual, off, K, 20)
    with dkb.KmerCounter(K, tuning=tuning) as kc:
        kc.build_table(entries)
        kc.submit_device(b2.data_ptr(), m1.data_ptr(), n_pos, 0)
        got = kc.entry_counts()[0]
        with pytest.raises(dkb.DkbError):  # one position more than the limit
            kc.submit_device(b2.data_ptr(), m1.data_ptr(), n_max + 1, 0)
    assert np.array_equal(got.astype(np.uint64), want), tuning
    assert want.sum() > 0
