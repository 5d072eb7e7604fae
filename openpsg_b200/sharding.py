"""Image sharding across the GPUs of one box.

The head is batch-1 by construction (``relation_transformer_head_v4.py:112``) and images share no state, so the
multi-GPU path is a partition of the image list: no collective on the data path, one gather of the small
per-image result records at the end.  Works over any initialised ``torch.distributed`` backend (NCCL on the GPU
box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import torch.distributed as dist


def shard_indices(num_items: int, rank: int, world: int) -> List[int]:
    """Round-robin: item i -> rank i % world (cfg4: 32 images -> 4 per GPU; cfg5: 8 images -> 1 per GPU)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, num_items, world))


def lpt_assign(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment for heterogeneous images (cost ~ N^2 pair queries):
    returns per-rank index lists; deterministic (ties -> lower index, lower rank)."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += float(costs[i])
    return [sorted(x) for x in out]


def gather_by_index(local: Dict[int, object], num_items: int) -> List[object]:
    """All ranks contribute {global index: result}; every rank gets the full list in index order."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts: List[Dict[int, object]] = [None] * dist.get_world_size()  # type: ignore[list-item]
        dist.all_gather_object(parts, local)
    else:
        parts = [local]
    merged: Dict[int, object] = {}
    for part in parts:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"item {k} produced by more than one rank")
            merged[k] = v
    missing = [i for i in range(num_items) if i not in merged]
    if missing:
        raise RuntimeError(f"items {missing[:8]} were not produced by any rank")
    return [merged[i] for i in range(num_items)]


def run_sharded(process: Callable[[int], object], num_items: int, costs: Sequence[float] | None = None) -> List[object]:
    """Run ``process(i)`` for this rank's share of ``range(num_items)`` and gather all results in order."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = lpt_assign(costs, world)[rank] if costs is not None else shard_indices(num_items, rank, world)
    return gather_by_index({i: process(i) for i in mine}, num_items)
