"""Multi-GPU plumbing: one process per GPU, read batches sharded across ranks, the
spanning-k-mer table replicated, and ONE sum-allreduce of the per-entry count vector
(north_star).  No other collective exists on the path.  CountPipeline overlaps
that allreduce with the next batch's scan.

The same functions run on `gloo` with CPU tensors (tests, world_size 2) and on `nccl`
with a tensor view of the library's device counters.
"""
import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(
        os.environ.get("LOCAL_RANK", 0))


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of n_items read batches (or reads) for rank."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevArray:
    """Expose a raw device pointer through __cuda_array_interface__ (int32 view)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False),
                                         "version": 2}


def counts_tensor(kc):
    """torch int32 view (no copy) of the context's [3][n_entries] device counters.
    Sums wrap mod 2^32 exactly like the uint32 counters do."""
    import torch
    ptr, n = kc.entry_counts_device()
    if n == 0:
        return torch.zeros(0, dtype=torch.int32, device=f"cuda:{kc.device}")
    return torch.as_tensor(_DevArray(ptr, n), device=f"cuda:{kc.device}")


def allreduce_counts(t, group=None):
    """In-place sum over ranks of a count tensor (device int32 view, or CPU int64/int32)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_counts_numpy(counts: np.ndarray, group=None) -> np.ndarray:
    """Host variant used by the gloo tests: uint32/uint64 array -> summed copy."""
    import torch
    t = torch.from_numpy(counts.astype(np.int64))
    allreduce_counts(t, group)
    return t.numpy().astype(counts.dtype)


class CountPipeline:
    """Overlap the allreduce of one batch's counters with the scan of the next batch.

    push() - call it when a batch's scans have been submitted - copies the context's counters
    into one of two buffers on the scan stream and starts the sum over ranks on a side
    stream; the PREVIOUS batch's reduced buffer is then finalised (kernel 3) on the scan
    stream, behind this batch's scans, by which time its allreduce has long finished.
    flush() finalises the last batch.  With world size 1 the allreduce is a no-op and the
    results equal plain finalise().
    """

    def __init__(self, kc, thresholds, group=None):
        import torch
        self.kc, self.thr, self.group = kc, thresholds, group
        self.dev = torch.device(f"cuda:{kc.device}")
        self.counts = counts_tensor(kc)
        self.bufs = [torch.empty_like(self.counts) for _ in range(2)]
        self.scan = torch.cuda.ExternalStream(kc.scan_stream(), device=self.dev)
        self.side = torch.cuda.Stream(device=self.dev)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.reduced = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self.pending = None  # buffer index whose finalise is still due

    def push(self):
        import torch
        b = self.i & 1
        self.i += 1
        with torch.cuda.stream(self.scan):
            self.bufs[b].copy_(self.counts, non_blocking=True)
            self.copied[b].record(self.scan)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.copied[b])
            allreduce_counts(self.bufs[b], self.group)
            self.reduced[b].record(self.side)
        prev, self.pending = self.pending, b
        return prev

    def finalise(self, b):
        if b is None:
            return
        self.scan.wait_event(self.reduced[b])
        self.kc.finalise_launch(self.thr, counts_ptr=self.bufs[b].data_ptr())

    def flush(self):
        b, self.pending = self.pending, None
        self.finalise(b)
