"""Scratch perf probe: scan a random device-resident stream against a table of n_variants
SNV candidates; prints time, Gbases/s and the stage counters for each tuning."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import denovo_kmer_b200 as dkb
from denovo_kmer_b200 import synth

n_var = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_pos = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 30
k = int(sys.argv[3]) if len(sys.argv) > 3 else 31
tunings = [tuple(int(x) for x in t.split(',')) for t in sys.argv[4:]] or [(15,1,1),(15,1,2),(15,2,1),(15,2,2),(14,4,1),(14,4,2)]
dev = torch.device('cuda:0')
genome = synth.make_genome(max(4_000_000, n_var * 400), 1)
variants = synth.plant_variants(genome, n_var, k, 2)
trio = synth.Trio(k, genome, variants)
entries = dkb.variant_kmers(trio.variant_tuples(), k)
g = torch.Generator(device=dev); g.manual_seed(3)
bw = (n_pos + 63)//64*4; mw = (n_pos+127)//128*4
bases = torch.randint(-2**31, 2**31-1, (bw,), dtype=torch.int32, device=dev, generator=g)
# separator every 151st position: approximate with all-valid mask except 1/151 random bits
mask = torch.full((mw,), -1, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for tun in tunings:
    for hints in (True, False):
        kc = dkb.KmerCounter(k, tuning=tun)
        kc.build_table(entries, use_window_hints=hints)
        st0 = kc.stats()
        for prof in (False, True):
            kc.profile_counters(prof)
            ms = []
            for it in range(4):
                kc.reset_counts()
                kc.submit_device(bases.data_ptr(), mask.data_ptr(), n_pos, 0)
                kc.sync()
                ms.append(kc.stats()['last_scan_ms'])
            s = kc.stats()
            if not prof:
                best = min(ms[1:])
                print(json.dumps(dict(tuning=tun, hints=hints, ms=round(best,3), gpos_s=round(n_pos/best/1e6,1),
                      n_entries=st0['n_entries'], n_seeds=st0['n_seeds'], bloom_density=round(st0['bloom_bits_set']/(st0['bloom_words']*32),4))))
            else:
                lookups = n_pos / tun[1]
                print('   prof ms=%.3f bloom_hit_rate=%.4f seed_hits=%d windows=%d hits=%d' % (min(ms[1:]), s['bloom_hits']/lookups, s['seed_hits'], s['windows_probed'], s['window_hits']))
        kc.close()
