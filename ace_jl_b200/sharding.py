"""Sharding of a ragged batch of environments over ranks (one process per GPU).

Environments are independent (`evaluate(model, cfg)` reads one configuration and immutable tables), so the
multi-GPU driver needs no data-path collective: each rank evaluates a contiguous range of environments,
balanced by neighbour count, and only the total energy is all-reduced.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_bounds(offsets, world: int) -> List[Tuple[int, int]]:
    """Contiguous environment ranges [(e0, e1)] per rank with near-equal neighbour counts."""
    offsets = np.asarray(offsets, dtype=np.int64)
    nenv = len(offsets) - 1
    total = int(offsets[-1] - offsets[0])
    cuts = [0]
    for r in range(1, world):
        target = offsets[0] + (total * r) // world
        e = int(np.searchsorted(offsets, target, side="left"))
        cuts.append(min(max(e, cuts[-1]), nenv))
    cuts.append(nenv)
    # every rank that can get an environment gets at least one
    for r in range(1, world):
        if cuts[r] <= cuts[r - 1] and cuts[r - 1] < nenv:
            cuts[r] = cuts[r - 1] + 1
    for r in range(world - 1, 0, -1):
        cuts[r] = min(cuts[r], cuts[r + 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def total_energy_allreduce(E_local, group=None):
    """Sum the per-environment energies of this rank and all-reduce over ranks (NCCL on GPU tensors)."""
    import torch
    import torch.distributed as dist
    tot = E_local.sum(dim=0) if isinstance(E_local, torch.Tensor) else torch.as_tensor(np.asarray(E_local).sum(axis=0))
    tot = tot.reshape(-1).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tot, group=group)
    return tot
