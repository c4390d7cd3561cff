"""Regenerates tests/golden/golden_small.json from the CPU oracle (oracle/).  These are
REGRESSION vectors of this repo's own spec: the reference's source and
tests are not mounted, so no reference-made golden vectors exist.
Run: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from denovo_kmer_b200 import synth  # noqa: E402


def case(k, seed, n_var, indel_frac, drop_shared, min_bq, thresholds):
    trio = synth.make_trio_host(4000, 5, n_var, k, seed=seed, read_len=60, indel_frac=indel_frac,
                                n_rate=0.01, lowq_frac=0.08)
    variants = trio.variant_tuples()
    keys, var, al, wi, wc = oracle.variant_entries(variants, k, drop_shared=drop_shared)
    ks = oracle.KmerSet(keys, var, al)
    counts = np.zeros((3, len(keys)), dtype=np.uint64)
    reads, quals = [], []
    for smp in range(3):
        seq, qual, off = trio.reads[smp]
        ks.count_reads(seq, qual, off, k, min_bq, counts=counts[smp])
        reads.append([seq[int(a):int(b)].tobytes().decode() for a, b in zip(off[:-1], off[1:])])
        quals.append([(qual[int(a):int(b)] + 33).tobytes().decode() for a, b in zip(off[:-1], off[1:])])
    hits, dist, nk = ks.variant_stats(counts, len(variants))
    return {"k": k, "min_bq": min_bq, "drop_shared": drop_shared, "thresholds": list(thresholds),
            "variants": [list(v) for v in variants], "reads": reads, "quals": quals,
            "entry_keys": keys.tolist(), "entry_variant": var.tolist(), "entry_allele": al.tolist(),
            "entry_counts": counts.tolist(), "hits": hits.tolist(), "distinct": dist.tolist(),
            "n_kmers": nk.tolist(), "calls": oracle.calls(hits, dist, thresholds).tolist()}


if __name__ == "__main__":
    cases = [case(31, 1, 4, 0.0, True, 20, (3, 2, 0, 1)),
             case(21, 2, 5, 0.5, True, 20, (3, 2, 0, 1)),
             case(15, 3, 5, 0.5, False, 10, (2, 1, 1, 1)),
             case(25, 4, 4, 1.0, True, 0, (3, 2, 0, 1))]
    with open(os.path.join(HERE, "golden_small.json"), "w") as f:
        json.dump({"note": "regression vectors from oracle/ (spec DESIGN.md §2); not reference outputs",
                   "cases": cases}, f, separators=(",", ":"))
    print("wrote", len(cases), "cases")
