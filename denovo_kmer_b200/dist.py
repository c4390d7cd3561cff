"""Multi-GPU plumbing: one process per GPU, read batches sharded across ranks, the
spanning-k-mer table replicated, and ONE sum-allreduce of the per-entry count vector
(north_star).  No other collective exists on the path.

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
