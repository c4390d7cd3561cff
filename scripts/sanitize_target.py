"""Small end-to-end run for compute-sanitizer: every tuning family once, parity vs oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import denovo_kmer_b200 as dkb, oracle
from denovo_kmer_b200 import synth
from helpers import gpu_counts, oracle_counts
trio = synth.make_trio_host(40_000, 8, 12, 31, seed=3, indel_frac=0.4, n_rate=0.01, lowq_frac=0.1)
entries = dkb.variant_kmers(trio.variant_tuples(), 31)
ks, want = oracle_counts(oracle, entries, trio, 31, 20)
for tun in [None, (15, 1, 1), (15, 2, 2), (14, 4, 2), (15, 8, 2), (15, 16, 1), (15, 16, 2, 2), (14, 4, 1, 2)]:
    for hints in (True, False):
        got, st, t = gpu_counts(dkb, entries, trio, 31, 20, tuning=tun, hints=hints, batches=3)
        assert np.array_equal(got.astype(np.uint64), want), (tun, hints)
print("sanitize target ok", int(want.sum()))
