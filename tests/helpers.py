"""Shared test helpers: run one trio through the oracle and through the CUDA path."""
import numpy as np


def oracle_counts(orc, entries, trio, k, min_bq):
    ks = orc.KmerSet(entries.keys, entries.variant, entries.allele)
    counts = np.zeros((3, len(entries)), dtype=np.uint64)
    for smp in range(3):
        seq, qual, off = trio.reads[smp]
        ks.count_reads(seq, qual, off, k, min_bq, counts=counts[smp])
    return ks, counts


def gpu_counts(dkb, entries, trio, k, min_bq, tuning=None, hints=True, batches=1, prof=False):
    with dkb.KmerCounter(k, tuning=tuning) as kc:
        kc.build_table(entries, use_window_hints=hints)
        if prof:
            kc.profile_counters(True)
        for smp in range(3):
            seq, qual, off = trio.reads[smp]
            n = len(off) - 1
            cuts = np.linspace(0, n, batches + 1).astype(int)
            for a, b in zip(cuts[:-1], cuts[1:]):
                if b == a:
                    continue
                lo, hi = int(off[a]), int(off[b])
                st = dkb.pack_reads(seq[lo:hi], None if qual is None else qual[lo:hi],
                                    off[a:b + 1] - off[a], min_bq)
                kc.submit(st, smp)
        counts = kc.entry_counts()
        stats = kc.stats()
        tun = kc.tuning()
    return counts, stats, tun
