"""CPU, world_size 2 over gloo: the multi-GPU plumbing (read-batch sharding + the single
sum-allreduce of the per-entry counters).  The per-rank scan is stood in by the oracle —
this tests the host logic around the kernel, not the kernel."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as tdist
    import oracle
    import denovo_kmer_b200 as dkb
    from denovo_kmer_b200 import dist, synth
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    assert dist.env_rank_world() == (rank, world, rank)
    k = 21
    trio = synth.make_trio_host(40_000, 10, 12, k, seed=31)
    entries = dkb.variant_kmers(trio.variant_tuples(), k)
    ks = oracle.KmerSet(entries.keys, entries.variant, entries.allele)
    counts = np.zeros((3, len(entries)), dtype=np.uint64)
    for smp in range(3):
        seq, qual, off = trio.reads[smp]
        lo, hi = dist.shard_range(len(off) - 1, rank, world)  # this rank's reads
        a, b = int(off[lo]), int(off[hi])
        ks.count_reads(seq[a:b], qual[a:b], off[lo:hi + 1] - off[lo], k, 20, counts=counts[smp])
    total = dist.allreduce_counts_numpy(counts)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), total)
    if rank == 0:
        full = np.zeros_like(counts)
        for smp in range(3):
            seq, qual, off = trio.reads[smp]
            ks.count_reads(seq, qual, off, k, 20, counts=full[smp])
        np.save(os.path.join(out_dir, "full.npy"), full)
    tdist.destroy_process_group()


def test_shard_range_partitions():
    from denovo_kmer_b200 import dist
    for n in (0, 1, 7, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharded_count_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = np.load(tmp_path / "full.npy")
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"rank{r}.npy"), full)
    assert full.sum() > 0
