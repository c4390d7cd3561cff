"""Measure scan throughput over (seed_len, stride, hashes) for a workload shape; prints a table.
Usage: python scripts/tuning_sweep.py K N_VARIANTS GENOME_MB [combos...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import denovo_kmer_b200 as dkb
from denovo_kmer_b200 import synth

k = int(sys.argv[1]); n_var = int(sys.argv[2]); gmb = float(sys.argv[3])
combos = [tuple(int(x) for x in c.split(',')) for c in sys.argv[4:]]
dev = torch.device('cuda:0')
genome = synth.make_genome(int(gmb * 1e6), 1)
variants = synth.plant_variants(genome, n_var, k, 2)
entries = dkb.variant_kmers(synth.Trio(k, genome, variants).variant_tuples(), k)
lut = np.zeros(256, dtype=np.uint8)
for i, ch in enumerate(b"ACGT"): lut[ch] = i
ref = torch.from_numpy(lut[genome]).to(dev)
alt = torch.from_numpy(lut[synth.apply_variants(genome, variants)]).to(dev)
n_reads = int(len(genome) * 30 / 150) // 128 * 128
b2, m1, n_pos, n_bases = synth.make_sample_device([ref, alt], n_reads, 150, 11, dev)
torch.cuda.synchronize()
base = None
for tun in [None] + combos:
    try:
        kc = dkb.KmerCounter(k, tuning=tun)
        kc.build_table(entries)
    except dkb.DkbError as e:
        print(tun, 'invalid:', str(e)[:60]); continue
    ms = []
    for it in range(4):
        kc.reset_counts(); kc.submit_device(b2.data_ptr(), m1.data_ptr(), n_pos, 0); kc.sync()
        ms.append(kc.stats()['last_scan_ms'])
    c = kc.entry_counts()[0]
    if base is None: base = c.copy()
    st = kc.stats()
    print(f"k={k} nvar={n_var} tuning={kc.tuning()} {'auto' if tun is None else ''} seeds={st['n_seeds']} "
          f"T/s={n_bases / min(ms[1:]) / 1e9:.3f} same_counts={bool(np.array_equal(c, base))}")
    kc.close()
