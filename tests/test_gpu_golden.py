"""-m gpu: the CUDA path against the committed golden fixtures (tests/golden/golden_small.json:
entries, per-entry counts, per-variant statistics and call bytes), without the oracle in
the loop."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("tuning", [None, (15, 1, 2), (13, 2, 1), (12, 4, 2), (10, 8, 1, 2)])
def test_golden_cases(dkb, tuning):
    with open(os.path.join(HERE, "golden", "golden_small.json")) as f:
        G = json.load(f)
    for case in G["cases"]:
        k = case["k"]
        if tuning is not None and tuning[0] > k - tuning[1] + 1:
            continue
        entries = dkb.variant_kmers([tuple(v) for v in case["variants"]], k,
                                    drop_shared=case["drop_shared"])
        assert entries.keys.tolist() == case["entry_keys"]
        assert entries.variant.tolist() == case["entry_variant"]
        assert entries.allele.tolist() == case["entry_allele"]
        with dkb.KmerCounter(k, tuning=tuning) as kc:
            kc.build_table(entries)
            for smp in range(3):
                seqs, quals = case["reads"][smp], case["quals"][smp]
                off = np.zeros(len(seqs) + 1, dtype=np.uint64)
                off[1:] = np.cumsum([len(s) for s in seqs])
                seq = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
                qual = np.frombuffer("".join(quals).encode(), dtype=np.uint8) - 33
                kc.submit(dkb.pack_reads(seq, qual, off, case["min_bq"]), smp)
            counts = kc.entry_counts()
            hits, dist, nk, calls = kc.finalise(case["thresholds"])
        assert counts.tolist() == case["entry_counts"], (k, tuning)
        assert hits.tolist() == case["hits"] and dist.tolist() == case["distinct"]
        assert nk.tolist() == case["n_kmers"] and calls.tolist() == case["calls"]
