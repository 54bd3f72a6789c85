"""Multi-GPU plumbing of the hot path: one process per GPU.

* Inference shards by window: rank r takes the contiguous slice ``shard_range(B, r, world)``; windows are
  independent, weights are replicated, there is NO collective on the data path.
* Training is data-parallel: every rank computes loss/gradients of its slice normalised by the GLOBAL
  config BATCH_SIZE (train.py:482, :488-489), so ranks combine gradients with a SUM all-reduce
  (``allreduce_gradients``); the AdamW update is then identical on every rank.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of n windows: the first n % world ranks get one extra window."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_gradients(flat_grad, dist) -> None:
    """Sum the flat gradient buffer over all ranks (NCCL on GPUs, gloo in the CPU tests)."""
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
