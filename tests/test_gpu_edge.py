"""-m gpu: edge cases of the CUDA path against the oracle — empty and degenerate inputs,
low-complexity sequence (dense filter hits: the multi-pass compaction path), keys shared
by many owners, repeated triples, maximum-length batches of tiny reads."""
import numpy as np
import pytest

import denovo_kmer_b200 as dkb_mod
from denovo_kmer_b200 import synth

pytestmark = pytest.mark.gpu


def _run(dkb, orc, entries, reads, k, min_bq=20, tuning=None, hints=True):
    """reads: list of 3 (seq, qual, offsets); returns (gpu counts, oracle counts)."""
    ks = orc.KmerSet(entries.keys, entries.variant, entries.allele)
    want = np.zeros((3, len(entries)), dtype=np.uint64)
    with dkb.KmerCounter(k, tuning=tuning) as kc:
        kc.build_table(entries, use_window_hints=hints)
        for smp, (seq, qual, off) in enumerate(reads):
            ks.count_reads(seq, qual, off, k, min_bq, counts=want[smp])
            kc.submit(dkb.pack_reads(seq, qual, off, min_bq), smp)
        got = kc.entry_counts()
        res = kc.finalise(dkb.DEFAULT_THRESHOLDS)
    return got, want, ks, res


def _ascii(s):
    return np.frombuffer(s.encode(), dtype=np.uint8)


def test_empty_inputs(dkb, orc):
    ent = dkb.variant_kmers([("ACGTACGTACGTACGTACGTACGTACGTAC", "A", "G", "TTGACCATGACCATTGACCATGACCAGGTT")], 31)
    empty = (np.zeros(0, np.uint8), None, np.zeros(1, np.uint64))
    got, want, _, _ = _run(dkb, orc, ent, [empty, empty, empty], 31)
    assert got.sum() == 0 and want.sum() == 0
    # empty table, non-empty reads
    none = dkb.KmerEntries(np.zeros(0, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint8),
                           None, None, 0)
    seq = _ascii("ACGT" * 50)
    r = (seq, None, np.array([0, len(seq)], dtype=np.uint64))
    with dkb.KmerCounter(31) as kc:
        kc.build_table(none)
        kc.submit(dkb.pack_reads(*r, 20), 0)
        assert kc.entry_counts().shape == (3, 0)
        hits, dist, nk, calls = kc.finalise()
        assert len(calls) == 0


def test_reads_shorter_than_k_and_all_masked(dkb, orc):
    trio = synth.make_trio_host(20_000, 10, 5, 31, seed=41)
    ent = dkb.variant_kmers(trio.variant_tuples(), 31)
    seq, qual, off = trio.reads[0]
    # (a) every read cut to 30 bases: no 31-mer exists
    n = len(off) - 1
    short_off = np.arange(0, 30 * (n + 1), 30, dtype=np.uint64)
    short_seq = seq.reshape(n, 150)[:, :30].reshape(-1).copy()
    short_q = qual.reshape(n, 150)[:, :30].reshape(-1).copy()
    # (b) all qualities below the threshold
    low_q = np.full_like(qual, 5)
    got, want, _, _ = _run(dkb, orc, ent, [(short_seq, short_q, short_off), (seq, low_q, off),
                                            (seq, qual, off)], 31)
    assert np.array_equal(got.astype(np.uint64), want)
    assert got[0].sum() == 0 and got[1].sum() == 0 and got[2].sum() > 0


@pytest.mark.parametrize("tuning", [None, (15, 1, 1), (15, 2, 2), (14, 4, 2), (8, 4, 1), (12, 8, 2), (12, 4, 2, 2)])
def test_low_complexity_dense_hits(dkb, orc, tuning):
    """Poly-A / dinucleotide / short tandem repeats: the same seed recurs at every position,
    filter hits exceed the per-tile id list, keys repeat inside one allele."""
    k = 21
    rng = np.random.default_rng(7)
    unit = lambda u, n: (u * (n // len(u) + 1))[:n]
    variants = [
        (unit("A", 20), "A", "C", unit("A", 20)),            # poly-A flanks
        (unit("AC", 20), "A", "G", unit("CA", 20)),           # dinucleotide repeat
        (unit("ACG", 20), "ACG", "A", unit("ACG", 20)),       # tandem-repeat deletion
        (unit("T", 20), "T", "TTT", unit("T", 20)),           # homopolymer insertion
        ("".join("ACGT"[i] for i in rng.integers(0, 4, 20)), "G", "T",
         "".join("ACGT"[i] for i in rng.integers(0, 4, 20))),
    ]
    for drop in (True, False):
        ent = dkb.variant_kmers(variants, k, drop_shared=drop)
        reads = []
        for smp in range(3):
            parts = [unit("A", 3000), unit("AC", 3000), unit("ACG", 3000), unit("T", 2000) + "C" + unit("A", 500),
                     variants[4][0] + "T" + variants[4][3], unit("CA", 777), unit("GT", 1500)]
            seqs = [p for p in parts for _ in range(3 + smp)]
            off = np.zeros(len(seqs) + 1, dtype=np.uint64)
            off[1:] = np.cumsum([len(s) for s in seqs])
            reads.append((_ascii("".join(seqs)), None, off))
        got, want, _, _ = _run(dkb, orc, ent, reads, k, tuning=tuning, hints=True)
        assert np.array_equal(got.astype(np.uint64), want), np.nonzero(got != want)
        got, want, _, _ = _run(dkb, orc, ent, reads, k, tuning=tuning, hints=False)
        assert np.array_equal(got.astype(np.uint64), want)
        assert want.sum() > 1000


def test_shared_keys_and_repeated_triples(dkb, orc):
    """One key owned by several (variant, allele) pairs; exact triples repeated."""
    k = 25
    trio = synth.make_trio_host(30_000, 12, 8, k, seed=43)
    ent = dkb.variant_kmers(trio.variant_tuples(), k)
    # every entry duplicated under two more variant ids + an exact repeat of the first 40
    nv = ent.n_variants
    keys = np.concatenate([ent.keys, ent.keys, ent.keys, ent.keys[:40]])
    var = np.concatenate([ent.variant, ent.variant + nv, (ent.variant + 1) % nv + 2 * nv,
                          ent.variant[:40]])
    al = np.concatenate([ent.allele, ent.allele, 1 - ent.allele, ent.allele[:40]])
    multi = dkb.KmerEntries(keys, var, al, None, None, 3 * nv)
    got, want, ks, res = _run(dkb, orc, multi, [trio.reads[s] for s in range(3)], k)
    assert np.array_equal(got.astype(np.uint64), want)
    n = len(ent)
    assert np.array_equal(got[:, :n], got[:, n:2 * n]) and got[:, 3 * n:].sum() == 0
    live = ks.live()
    assert live[: 3 * n].all() and not live[3 * n:].any()
    hits, dist, nk, calls = res
    o_hits, o_dist, o_nk = ks.variant_stats(want, 3 * nv)
    assert np.array_equal(hits.astype(np.uint64), o_hits) and np.array_equal(dist.astype(np.uint64), o_dist)
    assert np.array_equal(nk, o_nk)
    assert np.array_equal(calls, orc.calls(o_hits, o_dist, dkb.DEFAULT_THRESHOLDS))


def test_many_tiny_reads_and_batch_boundaries(dkb, orc):
    """Reads of length exactly k and k+1, submitted in batches of odd sizes: windows must
    never span a read separator or a batch boundary."""
    k = 15
    trio = synth.make_trio_host(10_000, 5, 6, k, seed=45)
    ent = dkb.variant_kmers(trio.variant_tuples(), k)
    g = trio.genome
    rng = np.random.default_rng(3)
    starts = rng.integers(0, len(g) - 20, size=40_000)
    lens = rng.integers(k, k + 2, size=len(starts))
    off = np.zeros(len(starts) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    seq = np.concatenate([g[s:s + l] for s, l in zip(starts, lens)])
    ks = orc.KmerSet(ent.keys, ent.variant, ent.allele)
    want = ks.count_reads(seq, None, off, k, 0)
    with dkb.KmerCounter(k) as kc:
        kc.build_table(ent)
        cuts = [0, 1, 2, 129, 130, 5000, 5001, 39_999, 40_000]
        for a, b in zip(cuts[:-1], cuts[1:]):
            lo, hi = int(off[a]), int(off[b])
            kc.submit(dkb.pack_reads(seq[lo:hi], None, off[a:b + 1] - off[a], 0), 0)
        got = kc.entry_counts()[0]
    assert np.array_equal(got.astype(np.uint64), want) and want.sum() > 0


def test_state_errors(dkb):
    from denovo_kmer_b200 import _lib
    with dkb.KmerCounter(31) as kc:
        st = dkb.pack_reads(_ascii("ACGT" * 20), None, np.array([0, 80], dtype=np.uint64), 0)
        with pytest.raises(dkb.DkbError) as ei:
            kc.submit(st, 0)
        assert ei.value.code == _lib.ESTATE
        ent = dkb.variant_kmers([("ACGTACGTACGTACGTACGTACGTACGTAC", "A", "G", "TTGACCATGACCATTGACCATGACCAGGTT")], 31)
        kc.build_table(ent)
        with pytest.raises(dkb.DkbError) as ei:
            kc.submit(st, 3)
        assert ei.value.code == _lib.EINVAL
        with pytest.raises(dkb.DkbError):
            kc.results()  # finalise has not run
        with pytest.raises(dkb.DkbError):
            kc.set_tuning(16, 1, 1) or kc.build_table(ent)  # seed_len 16 is outside 8..15


def test_context_lifecycle(dkb, orc):
    """Rebuild the table inside one context, run two contexts side by side, reset counters:
    each result must still equal the oracle's."""
    trio_a = synth.make_trio_host(30_000, 10, 8, 31, seed=51)
    trio_b = synth.make_trio_host(30_000, 10, 8, 21, seed=52, indel_frac=0.5)
    ent_a = dkb.variant_kmers(trio_a.variant_tuples(), 31)
    ent_b31 = dkb.variant_kmers(trio_b.variant_tuples(), 31)
    ent_b = dkb.variant_kmers(trio_b.variant_tuples(), 21)

    def want(ent, trio, k):
        ks = orc.KmerSet(ent.keys, ent.variant, ent.allele)
        return ks.count_reads(*trio.reads[0], k, 20)

    with dkb.KmerCounter(31) as k1, dkb.KmerCounter(21) as k2:
        k1.build_table(ent_a)
        k2.build_table(ent_b)
        k1.submit(dkb.pack_reads(*trio_a.reads[0], 20), 0)
        k2.submit(dkb.pack_reads(*trio_b.reads[0], 20), 0)
        assert np.array_equal(k1.entry_counts()[0].astype(np.uint64), want(ent_a, trio_a, 31))
        assert np.array_equal(k2.entry_counts()[0].astype(np.uint64), want(ent_b, trio_b, 21))
        # counters accumulate across submits until reset
        k1.submit(dkb.pack_reads(*trio_a.reads[0], 20), 0)
        assert np.array_equal(k1.entry_counts()[0].astype(np.uint64), 2 * want(ent_a, trio_a, 31))
        k1.reset_counts()
        assert k1.entry_counts().sum() == 0
        # a second build in the same context replaces table, seeds, filter and counters
        k1.set_tuning(15, 2, 2)
        k1.build_table(ent_b31)
        k1.submit(dkb.pack_reads(*trio_b.reads[0], 20), 0)
        assert np.array_equal(k1.entry_counts()[0].astype(np.uint64), want(ent_b31, trio_b, 31))
        assert k1.tuning()[:3] == (15, 2, 2)
        st = k1.stats()
        assert st["n_entries"] == len(ent_b31) and st["scan_launches"] == 3


def _to_bam4(seq, off):
    """ASCII reads -> BAM 4-bit codes, every read byte-aligned (what htslib keeps in memory)."""
    code = np.full(256, 15, np.uint8)  # N
    for ch, v in zip(b"=ACMGRSVTWYHKDBN", range(16)):
        code[ch] = v
        code[ord(chr(ch).lower())] = v
    out = []
    for a, b in zip(off[:-1], off[1:]):
        c = code[seq[int(a):int(b)]]
        if len(c) % 2:
            c = np.concatenate([c, np.zeros(1, np.uint8)])
        out.append((c[0::2] << 4) | c[1::2])
    return np.concatenate(out) if out else np.zeros(0, np.uint8)


@pytest.mark.parametrize("ragged", [False, True])
def test_device_packer_matches_host_packer(dkb, orc, ragged):
    """dkb_batch_submit_reads (packing on the GPU from ASCII or BAM 4-bit reads) must give
    the counts of the host-packed path and of the oracle."""
    k = 21
    trio = synth.make_trio_host(60_000, 12, 10, k, seed=61, ragged=ragged, n_rate=0.01,
                                lowq_frac=0.1, indel_frac=0.3)
    ent = dkb.variant_kmers(trio.variant_tuples(), k)
    ks = orc.KmerSet(ent.keys, ent.variant, ent.allele)
    want = np.zeros((3, len(ent)), dtype=np.uint64)
    for smp in range(3):
        ks.count_reads(*trio.reads[smp], k, 20, counts=want[smp])
    for four_bit in (False, True):
        with dkb.KmerCounter(k) as kc:
            kc.build_table(ent)
            for smp in range(3):
                seq, qual, off = trio.reads[smp]
                n = len(off) - 1
                cuts = [0, n // 3, n // 3 + 1, n]  # three batches, the middle one a single read
                for a, b in zip(cuts[:-1], cuts[1:]):
                    o = off[a:b + 1]
                    if four_bit:
                        kc.submit_reads(_to_bam4(seq, o), qual, o, smp, 20, four_bit=True)
                    else:
                        kc.submit_reads(seq, qual, o, smp, 20)
            got = kc.entry_counts()
        assert np.array_equal(got.astype(np.uint64), want), four_bit
    # no qualities, min_baseq irrelevant
    with dkb.KmerCounter(k) as kc:
        kc.build_table(ent)
        seq, qual, off = trio.reads[0]
        kc.submit_reads(seq, None, off, 0, 99)
        got = kc.entry_counts()[0]
    assert np.array_equal(got.astype(np.uint64), ks.count_reads(seq, None, off, k, 0))


def test_rebuild_flips_the_prefilter_within_one_context(dkb, orc, monkeypatch):
    """The dynamic shared-memory size of an L2-filter kernel follows the pre-filter every
    table build chooses.  Same kernel (stride, bits), pre-filter off -> on -> off inside one
    context, and a second context with the other choice alongside: every launch must get the
    size it needs and every result must equal the oracle's."""
    trio = synth.make_trio_host(60_000, 10, 12, 31, seed=61)
    ent = dkb.variant_kmers(trio.variant_tuples(), 31)
    ks = orc.KmerSet(ent.keys, ent.variant, ent.allele)
    want = ks.count_reads(*trio.reads[0], 31, 20)
    st = dkb.pack_reads(*trio.reads[0], 20)

    def run(kc, words):
        monkeypatch.setenv("DKB_PREFILTER_WORDS", str(words))
        kc.set_tuning(15, 16, 2, 2)
        kc.build_table(ent)
        assert kc.stats()["prefilter_words"] == words
        kc.submit(st, 0)
        assert np.array_equal(kc.entry_counts()[0].astype(np.uint64), want), words

    with dkb.KmerCounter(31) as k1, dkb.KmerCounter(31) as k2:
        for words in (0, 35584, 0, 51712, 4096):
            run(k1, words)
            run(k2, 35584 if words == 0 else 0)
            k1.reset_counts()
            k1.submit(st, 0)  # k1's kernel again after k2 changed the function's attributes
            assert np.array_equal(k1.entry_counts()[0].astype(np.uint64), want), words
    assert want.sum() > 0


@pytest.mark.parametrize("tuning", [None, (14, 4, 2, 1), (15, 16, 2, 1), (15, 16, 2, 2), (15, 8, 1, 2), (15, 1, 2, 1)])
def test_one_launch_for_a_trio_equals_three_launches(dkb, orc, tuning):
    """dkb_batch_submit_device_multi: the three samples' streams (different lengths, one of them
    empty in a second round) scanned by ONE launch give the counters of three launches."""
    import torch
    dev = torch.device("cuda:0")
    trio = synth.make_trio_host(150_000, 12, 30, 31, seed=71, indel_frac=0.3)
    ent = dkb.variant_kmers(trio.variant_tuples(), 31)
    ks = orc.KmerSet(ent.keys, ent.variant, ent.allele)
    want = np.zeros((3, len(ent)), dtype=np.uint64)
    dstreams = []
    for smp in range(3):
        seq, qual, off = trio.reads[smp]
        n = (len(off) - 1) * (smp + 1) // 3          # three different lengths
        off = off[:n + 1]
        ks.count_reads(seq, qual, off, 31, 20, counts=want[smp])
        st = dkb.pack_reads(seq[:int(off[-1])], qual[:int(off[-1])], off, 20)
        dstreams.append((torch.from_numpy(st.bases2.view(np.int32)).to(dev),
                         torch.from_numpy(st.mask1.view(np.int32)).to(dev), st.n_positions))
    with dkb.KmerCounter(31, tuning=tuning) as kc:
        kc.build_table(ent)
        kc.submit_device_multi([(b.data_ptr(), m.data_ptr(), n, s) for s, (b, m, n) in enumerate(dstreams)])
        assert kc.stats()["scan_launches"] == 1
        assert np.array_equal(kc.entry_counts().astype(np.uint64), want), tuning
        # same stream twice into one sample, an empty batch in between, sample order reversed
        kc.reset_counts()
        b, m, n = dstreams[1]
        kc.submit_device_multi([(b.data_ptr(), m.data_ptr(), n, 2), (b.data_ptr(), m.data_ptr(), 0, 1),
                                (b.data_ptr(), m.data_ptr(), n, 2), (dstreams[0][0].data_ptr(), dstreams[0][1].data_ptr(), dstreams[0][2], 0)])
        got = kc.entry_counts().astype(np.uint64)
        assert np.array_equal(got[2], 2 * want[1]) and np.array_equal(got[0], want[0]) and got[1].sum() == 0
        with pytest.raises(dkb.DkbError):
            kc.submit_device_multi([(b.data_ptr(), m.data_ptr(), n, 0)] * 5)
    assert want.sum() > 0


def test_padding_bits_beyond_the_stream_are_ignored(dkb, orc):
    """dkb_batch_submit_device takes caller-owned buffers: flag bits at positions >= n_positions
    (the padding of the last words) must not produce counts, whatever they hold."""
    import torch
    dev = torch.device("cuda:0")
    k = 31
    trio = synth.make_trio_host(40_000, 10, 10, k, seed=81, n_rate=0.0, lowq_frac=0.0, err_rate=0.0)
    ent = dkb.variant_kmers(trio.variant_tuples(), k)
    ks = orc.KmerSet(ent.keys, ent.variant, ent.allele)
    seq, qual, off = trio.reads[0]
    full = dkb.pack_reads(seq, qual, off, 20)
    for cut_reads in (len(off) - 1 - 1, (len(off) - 1) // 2):
        # scan only the first cut_reads reads, minus a few positions so the cut falls inside a
        # read: everything after the cut in the buffers is real, valid sequence = worst-case padding
        n_pos = cut_reads * 151 - 37
        want = np.zeros(len(ent), dtype=np.uint64)
        ks.count_reads(seq, qual, off[:cut_reads], k, 20, counts=want)   # whole reads before the cut one
        a = int(off[cut_reads - 1])
        tail = np.array([0, 151 - 37], dtype=np.uint64)
        ks.count_reads(seq[a:a + 150], qual[a:a + 150], tail, k, 20, counts=want)  # the cut read's first 114 bases
        b = torch.from_numpy(full.bases2.view(np.int32)).to(dev)
        m = torch.from_numpy(full.mask1.view(np.int32)).to(dev)
        for tuning in (None, (15, 16, 2, 2), (14, 4, 2, 1), (15, 1, 1, 1)):
            with dkb.KmerCounter(k, tuning=tuning) as kc:
                kc.build_table(ent)
                kc.submit_device(b.data_ptr(), m.data_ptr(), n_pos, 0)
                assert np.array_equal(kc.entry_counts()[0].astype(np.uint64), want), (cut_reads, tuning)
    assert want.sum() > 0


@pytest.mark.parametrize("tuning", [None, (15, 16, 2, 2), (14, 4, 2, 1)])
def test_zero_list_submit_equals_dense_submit(dkb, orc, tuning):
    """dkb_batch_submit_sparse (flags as a zero list, expanded on the device) gives the counters
    of the dense form and of the oracle; batches that end inside a block, a read made of N only
    (a plain-bits block) and an empty zero list included."""
    k = 31
    trio = synth.make_trio_host(120_000, 12, 25, k, seed=93, n_rate=0.004, lowq_frac=0.06)
    ent = dkb.variant_kmers(trio.variant_tuples(), k)
    ks = orc.KmerSet(ent.keys, ent.variant, ent.allele)
    want = np.zeros((3, len(ent)), dtype=np.uint64)
    with dkb.KmerCounter(k, tuning=tuning) as kc:
        kc.build_table(ent)
        for smp in range(3):
            seq, qual, off = trio.reads[smp]
            seq = seq.copy()
            if smp == 1:
                seq[int(off[10]):int(off[30])] = ord("N")   # twenty reads of N: dense zeros
            if smp == 2:
                qual = None                                  # hardly any zeros besides the separators
                seq[seq == ord("N")] = ord("A")
            ks.count_reads(seq, qual, off, k, 20, counts=want[smp])
            cuts = [0, 7, (len(off) - 1) // 2, len(off) - 1]
            for a, b in zip(cuts[:-1], cuts[1:]):
                lo, hi = int(off[a]), int(off[b])
                st = dkb.pack_reads(seq[lo:hi], None if qual is None else qual[lo:hi], off[a:b + 1] - off[a], 20)
                zoff, zbytes = dkb.mask_to_zero_list(st.mask1, st.n_positions)
                kc.submit_sparse(st.bases2, zoff, zbytes, st.n_positions, smp)
        got = kc.entry_counts()
    assert np.array_equal(got.astype(np.uint64), want), tuning
    assert want.sum() > 0


def test_malformed_zero_list_is_refused(dkb):
    """dkb_batch_submit_sparse checks the block offsets before the expander kernel sees them."""
    trio = synth.make_trio_host(30_000, 10, 5, 31, seed=9)
    ent = dkb.variant_kmers(trio.variant_tuples(), 31)
    st = dkb.pack_reads(*trio.reads[0], 20)
    zoff, zbytes = dkb.mask_to_zero_list(st.mask1, st.n_positions)
    with dkb.KmerCounter(31) as kc:
        kc.build_table(ent)
        kc.submit_sparse(st.bases2, zoff, zbytes, st.n_positions, 0)  # the well-formed list passes
        for mutate in (lambda z: z.__setitem__(1, int(z[2]) + 1),          # offsets descend
                       lambda z: z.__setitem__(len(z) // 2, len(zbytes) + 7),  # beyond the bytes
                       lambda z: z.__setitem__(0, int(z[0]) | 0x80000000)):   # "plain bits" block that is not 256 bytes
            bad = zoff.copy()
            mutate(bad)
            with pytest.raises(dkb.DkbError):
                kc.submit_sparse(st.bases2, bad, zbytes, st.n_positions, 0)
        kc.sync()
