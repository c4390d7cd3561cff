"""CPU: the C oracle against hand-derived known answers, a pure-Python restatement and
the committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py from
the oracle itself — regression vectors, NOT reference outputs: parity is unpinned, see
oracle/dnk_oracle.c)."""
import json
import os
import random

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_known_answers_kmer_primitives(orc):
    # ACG = 0b000110 = 6 ; rc(ACG) = CGT = 0b011011 = 27 ; canonical = 6
    assert orc.py_encode("ACG") == 6 and orc.revcomp(6, 3) == 27 and orc.canonical(27, 3) == 6
    # palindrome ACGT (k=4): rc == itself
    assert orc.revcomp(orc.py_encode("ACGT"), 4) == orc.py_encode("ACGT")
    # 31-mer of all T <-> all A
    assert orc.revcomp((1 << 62) - 1, 31) == 0 and orc.canonical((1 << 62) - 1, 31) == 0
    assert orc.lib().orc_base_code(ord("n")) == -1 and orc.lib().orc_base_code(ord("g")) == 2


def test_known_answer_read_kmers(orc):
    # hand-derived: k=3 over "ACGTNACGTAC": windows 0,1 then (N resets) 5..8
    keys, pos = orc.read_kmers(b"ACGTNACGTAC", None, 3)
    assert pos.tolist() == [0, 1, 5, 6, 7, 8]
    # ACG->6, CGT->rc ACG->6, ACG->6, CGT->6, GTA->min(GTA=44, TAC=49)=44, TAC->min(49, GTA=44)=44
    assert keys.tolist() == [6, 6, 6, 6, 44, 44]
    # a low-quality base at index 2 kills every window covering it
    q = np.array([30, 30, 5, 30, 30, 30], dtype=np.uint8)
    keys, pos = orc.read_kmers(b"ACGTAC", q, 3, 20)
    assert pos.tolist() == [3]
    keys, pos = orc.read_kmers(b"ACGTAC", q, 3, 5)  # threshold is >=
    assert pos.tolist() == [0, 1, 2, 3]
    assert len(orc.read_kmers(b"AC", None, 3)[0]) == 0 and len(orc.read_kmers(b"", None, 3)[0]) == 0


def test_known_answer_counts_and_calls(orc):
    # one SNV: left=AC, ref=G, alt=T, right=TA, k=3 -> ref hap ACGTA, alt hap ACTTA
    keys, var, al, wi, wc = orc.variant_entries([("AC", "G", "T", "TA")], 3)
    enc = orc.py_canonical_str
    assert keys[al == 0].tolist() == [enc("ACG"), enc("GTA")]        # CGT == rc(ACG): repeat dropped
    assert keys[al == 1].tolist() == [enc("ACT"), enc("CTT"), enc("TTA")]
    ks = orc.KmerSet(keys, var, al)
    seq = np.frombuffer(b"ACTTAACGTA", dtype=np.uint8)
    off = np.array([0, 5, 10], dtype=np.uint64)
    c = ks.count_reads(seq, None, off, 3)
    # read1 ACTTA: ACT, CTT, TTA -> alt 1,1,1 ; read2 ACGTA: ACG, CGT(=ACG), GTA -> ref ACG 2, GTA 1
    assert c.tolist() == [2, 1, 1, 1, 1]
    counts3 = np.zeros((3, 5), dtype=np.uint64)
    counts3[0] = c
    counts3[1] = [4, 4, 0, 0, 0]
    counts3[2] = [3, 3, 0, 0, 0]
    hits, dist, nk = ks.variant_stats(counts3, 1)
    assert hits[0].tolist() == [[3, 8, 6], [3, 0, 0]] and dist[0].tolist() == [[2, 2, 2], [3, 0, 0]]
    assert nk[0].tolist() == [2, 3]
    assert orc.calls(hits, dist, (3, 2, 0, 1)).tolist() == [0x01]
    assert orc.calls(hits, dist, (4, 2, 0, 1)).tolist() == [0x02]
    counts3[1, 2] = 1  # mother carries an alt k-mer
    hits, dist, nk = ks.variant_stats(counts3, 1)
    assert orc.calls(hits, dist, (3, 2, 0, 1)).tolist() == [0x04]
    assert orc.calls(hits, dist, (3, 2, 1, 9)).tolist() == [0x10]


def test_allele_kmers_indels(orc):
    k = 4
    # deletion REF=GAT ALT=G : alt hap = L[-3:] + G + R[:3]
    keys, win, run = orc.allele_kmers("TTACC", "G", "CATGG", k)
    hap = "ACC" + "G" + "CAT"
    assert run == 4 and keys.tolist() == [orc.py_canonical_str(hap[w:w + k]) for w in range(4)]
    # empty allele: only junction-straddling windows
    keys, win, run = orc.allele_kmers("TTACC", "", "CATGG", k)
    hap = "ACC" + "CAT"
    assert run == 3 and keys.tolist() == [orc.py_canonical_str(hap[w:w + k]) for w in range(3)]
    # N in a flank drops the windows that cover it; short flanks shorten the run
    keys, win, run = orc.allele_kmers("ANC", "G", "CA", k)
    assert run == 3 and win.tolist() == [2]


@pytest.mark.parametrize("k", [3, 5, 8, 15, 21, 31])
def test_c_oracle_vs_python_restatement(orc, k):
    rnd = random.Random(100 + k)
    alpha = "ACGT" * 8 + "Nacgt"
    entries = []
    reads = []
    for _ in range(40):
        n = rnd.randint(0, 3 * k + 5)
        s = "".join(rnd.choice(alpha) for _ in range(n))
        q = [rnd.choice([2, 10, 19, 20, 21, 35, 40]) for _ in range(n)] if rnd.random() < 0.7 else None
        reads.append((s, q))
    # entries: keys drawn from the reads' own k-mers, owners with repeats and shared keys
    pool = [key for s, q in reads for _, key in orc.py_read_kmers(s, None, k)]
    for _ in range(60):
        if pool:
            entries.append((rnd.choice(pool), rnd.randint(0, 4), rnd.randint(0, 1)))
    want = orc.py_count(entries, reads, k, 20)
    ks = orc.KmerSet([e[0] for e in entries], [e[1] for e in entries], [e[2] for e in entries])
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s, _ in reads])
    seq = np.frombuffer("".join(s for s, _ in reads).encode(), dtype=np.uint8)
    qual = np.concatenate([np.array(q if q is not None else [40] * len(s), dtype=np.uint8)
                           for s, q in reads]) if len(seq) else np.zeros(0, dtype=np.uint8)
    for threads in (1, 3):
        got = ks.count_reads(seq, qual, off, k, 20, threads=threads)
        assert got.tolist() == want
    for s, q in reads[:10]:
        keys, pos = orc.read_kmers(s.encode(), None if q is None else np.array(q, dtype=np.uint8), k, 20)
        assert list(zip(pos.tolist(), keys.tolist())) == orc.py_read_kmers(s, q, k, 20)


def test_golden_fixtures(orc):
    with open(os.path.join(HERE, "golden", "golden_small.json")) as f:
        G = json.load(f)
    for case in G["cases"]:
        k = case["k"]
        keys, var, al, wi, wc = orc.variant_entries([tuple(v) for v in case["variants"]], k,
                                                    drop_shared=case["drop_shared"])
        assert keys.tolist() == case["entry_keys"] and var.tolist() == case["entry_variant"]
        assert al.tolist() == case["entry_allele"]
        ks = orc.KmerSet(keys, var, al)
        counts = np.zeros((3, len(keys)), dtype=np.uint64)
        for smp in range(3):
            seqs = case["reads"][smp]
            quals = case["quals"][smp]
            off = np.zeros(len(seqs) + 1, dtype=np.uint64)
            off[1:] = np.cumsum([len(s) for s in seqs])
            seq = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
            qual = np.frombuffer("".join(quals).encode(), dtype=np.uint8) - 33  # phred+33 strings
            ks.count_reads(seq, qual, off, k, case["min_bq"], counts=counts[smp])
        assert counts.tolist() == case["entry_counts"]
        hits, dist, nk = ks.variant_stats(counts, len(case["variants"]))
        assert hits.tolist() == case["hits"] and dist.tolist() == case["distinct"]
        assert orc.calls(hits, dist, case["thresholds"]).tolist() == case["calls"]


def test_count_stream_equals_count_reads(orc):
    """The packed-stream form of the count (what the full-size GPU parity tests and bench.py
    use as the checker) equals the per-read form on ragged reads with N and low-quality
    bases, for several k and thread counts; block boundaries (2^20 positions) are crossed."""
    from denovo_kmer_b200 import synth
    for k, seed in ((31, 5), (15, 6), (8, 7)):
        trio = synth.make_trio_host(60_000, 40, 25, k, seed=seed, indel_frac=0.3, ragged=(k == 15),
                                    n_rate=0.003, lowq_frac=0.05)
        keys, var, al, _, _ = orc.variant_entries(trio.variant_tuples(), k)
        ks = orc.KmerSet(keys, var, al)
        for smp in range(3):
            seq, qual, off = trio.reads[smp]
            want = ks.count_reads(seq, qual, off, k, 20)
            b2, m1, n_pos = orc.pack_stream(seq, qual, off, 20)
            assert n_pos == len(seq) + len(off) - 1
            if k != 15:
                assert n_pos > (1 << 20)
            for threads in (1, 3):
                got = ks.count_stream(b2, m1, n_pos, k, threads=threads)
                assert np.array_equal(got, want), (k, smp, threads)
            assert want.sum() > 0


def test_pack_stream_known_answer(orc):
    # reads "ACGN" (q 30,30,5,30) and "TT": positions A C G N | T T |  -> 8 positions
    seq = np.frombuffer(b"ACGNTT", dtype=np.uint8)
    qual = np.array([30, 30, 5, 30, 30, 30], dtype=np.uint8)
    off = np.array([0, 4, 6], dtype=np.uint64)
    b2, m1, n_pos = orc.pack_stream(seq, qual, off, 20)
    assert n_pos == 8 and len(b2) == 4 and len(m1) == 4
    # codes: A=0 C=1 (G masked -> 0) (N -> 0) sep T=3 T=3 sep ; flags 1 1 0 0 0 1 1 0
    assert int(b2[0]) == (1 << 2) | (3 << 10) | (3 << 12) and int(m1[0]) == 0b01100011
