"""Independent MPC instances shard across GPUs with no collective on the solve path
(SURVEY.md §8(e)); the only exchange is a final gather of per-rank results."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block of instance indices owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def gather_summaries(local: torch.Tensor) -> List[torch.Tensor]:
    """Final gather: every rank contributes one small tensor (same shape); returns the
    list on every rank.  Works on NCCL (cuda tensors) and gloo (cpu tensors)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local.contiguous())
    return out


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
