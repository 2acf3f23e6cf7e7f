"""Batch sharding for the MMA path.  The path shards over batch only (each sequence's attention is independent;
the reference has no TP/SP/CP): eval / prefill / decode run one process per GPU on a contiguous
slice of the batch with NO data-path collective; SFT uses torch DDP (NCCL gradient all-reduce over NVLink)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n items for `rank`; the first n % world ranks take one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    lo, hi = shard_range(tensors[0].shape[0], rank, world)
    return [t[lo:hi] for t in tensors]


def gather_batch(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Inverse of shard_batch for results (token ids / logits slices): concatenation in rank order.  Off the timed
    path; uses all_gather_object so ragged shards and CPU (gloo) / GPU (nccl) groups both work."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local.cpu())
    out = torch.cat(parts, dim=0)
    assert out.shape[0] == n_total, (out.shape, n_total)
    return out


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
