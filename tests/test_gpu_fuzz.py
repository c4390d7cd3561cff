"""-m gpu: seeded random cases through every degree of freedom at once - k, seed length,
stride, filter bits, filter mode, pre-filter size, gated lookups, window hints, drop_shared, indels, ragged
reads, N / low-quality rates, batch cuts, one launch per trio or per batch, dense or zero-list
flags - each against the oracle, bit for bit.  Catches interactions the fixed parity cases
(tests/test_gpu_parity.py) do not enumerate."""
import os

import numpy as np
import pytest
import torch

from denovo_kmer_b200 import synth

pytestmark = pytest.mark.gpu


def _case(rng):
    k = int(rng.choice([8, 11, 15, 16, 21, 25, 30, 31]))
    D = int(rng.choice([1, 2, 4, 8, 16]))
    while D > 1 and k - D + 1 < 8:
        D //= 2
    s = int(rng.integers(8, min(15, k - D + 1) + 1))
    mode = int(rng.choice([1, 2])) if D >= 2 else 1
    bits = int(rng.integers(1, 3)) if mode == 2 else int(rng.integers(1, 5))
    return dict(k=k, tuning=(s, D, bits, mode), pre=int(rng.choice([0, 256, 35584])),
                hints=bool(rng.integers(0, 2)), drop_shared=bool(rng.integers(0, 2)),
                indel=float(rng.choice([0.0, 0.5])), ragged=bool(rng.integers(0, 2)),
                n_rate=float(rng.choice([0.0, 0.002, 0.02])), lowq=float(rng.choice([0.0, 0.03, 0.2])),
                batches=int(rng.integers(1, 4)), how=str(rng.choice(["host", "sparse", "multi"])),
                auto=bool(rng.integers(0, 4) == 0), gate=str(rng.choice(["", "0", "1"])),
                seed=int(rng.integers(1, 1 << 30)), glen=60_000, depth=12, n_var=25, read_len=150, min_bq=20)


def _wide_case(rng):
    """Soak runs (case_seed >= 10 000): every k, more shapes of input."""
    c = _case(rng)
    k = int(rng.integers(8, 32))
    D = int(rng.choice([1, 2, 4, 8, 16]))
    while D > 1 and k - D + 1 < 8:
        D //= 2
    s = int(rng.integers(8, min(15, k - D + 1) + 1))
    mode = int(rng.choice([1, 2])) if D >= 2 else 1
    bits = int(rng.integers(1, 3)) if mode == 2 else int(rng.integers(1, 5))
    c.update(k=k, tuning=(s, D, bits, mode), glen=int(rng.choice([20_000, 60_000, 200_000])),
             depth=int(rng.choice([3, 12, 30])), n_var=int(rng.choice([1, 25, 120])),
             read_len=int(rng.choice([40, 100, 150, 251])), min_bq=int(rng.choice([0, 10, 20, 35])))
    return c


# 40 cases in the regular run; DKB_FUZZ_CASES / DKB_FUZZ_BASE widen or move the range for a soak
# run (profiles/README.md has the last one)
_CASES = range(int(os.environ.get("DKB_FUZZ_BASE", "0")),
               int(os.environ.get("DKB_FUZZ_BASE", "0")) + int(os.environ.get("DKB_FUZZ_CASES", "40")))


@pytest.mark.parametrize("case_seed", _CASES)
def test_random_case(dkb, orc, case_seed, monkeypatch):
    rng = np.random.default_rng(1000 + case_seed)
    c = _wide_case(rng) if case_seed >= 10_000 else _case(rng)
    k, min_bq = c["k"], c["min_bq"]
    monkeypatch.setenv("DKB_PREFILTER_WORDS", str(c["pre"]))
    if c["gate"]:
        monkeypatch.setenv("DKB_GATE", c["gate"])  # lookups gated by the flag stream: forced on / off
    trio = synth.make_trio_host(c["glen"], c["depth"], c["n_var"], k, seed=c["seed"], read_len=max(c["read_len"], k + 2),
                                indel_frac=c["indel"], ragged=c["ragged"], n_rate=c["n_rate"], lowq_frac=c["lowq"])
    entries = dkb.variant_kmers(trio.variant_tuples(), k, drop_shared=c["drop_shared"])
    ks = orc.KmerSet(entries.keys, entries.variant, entries.allele)
    want = np.zeros((3, len(entries)), dtype=np.uint64)
    dev = torch.device("cuda:0")
    with dkb.KmerCounter(k, tuning=None if c["auto"] else c["tuning"]) as kc:
        kc.build_table(entries, use_window_hints=c["hints"])
        keep, multi = [], []
        for smp in range(3):
            seq, qual, off = trio.reads[smp]
            ks.count_reads(seq, qual, off, k, min_bq, counts=want[smp])
            n = len(off) - 1
            cuts = np.linspace(0, n, c["batches"] + 1).astype(int)
            for a, b in zip(cuts[:-1], cuts[1:]):
                if b == a:
                    continue
                lo, hi = int(off[a]), int(off[b])
                st = dkb.pack_reads(seq[lo:hi], qual[lo:hi], off[a:b + 1] - off[a], min_bq)
                if c["how"] == "host":
                    kc.submit(st, smp)
                elif c["how"] == "sparse":
                    zoff, zbytes = dkb.mask_to_zero_list(st.mask1, st.n_positions)
                    kc.submit_sparse(st.bases2, zoff, zbytes, st.n_positions, smp)
                else:
                    d = (torch.from_numpy(st.bases2.view(np.int32)).to(dev), torch.from_numpy(st.mask1.view(np.int32)).to(dev))
                    keep.append(d)
                    multi.append((d[0].data_ptr(), d[1].data_ptr(), st.n_positions, smp))
        for i in range(0, len(multi), 4):
            kc.submit_device_multi(multi[i:i + 4])
        got = kc.entry_counts()
        hits, dist, nk, calls = kc.finalise((3, 2, 0, 1))
        tun = kc.tuning()
    bad = int((got.astype(np.uint64) != want).sum())
    assert bad == 0, f"{bad} counters differ; case {c}; resolved tuning {tun}"
    o_hits, o_dist, o_nk = ks.variant_stats(want, entries.n_variants)
    assert np.array_equal(hits.astype(np.uint64), o_hits) and np.array_equal(dist.astype(np.uint64), o_dist)
    assert np.array_equal(nk, o_nk) and np.array_equal(calls, orc.calls(o_hits, o_dist, (3, 2, 0, 1)))
