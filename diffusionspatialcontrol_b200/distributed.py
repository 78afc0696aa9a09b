"""Data-parallel seed-batch sharding (the only way the path shards).

The std in the hot path couples every score of ONE attention call, so a call is never split across
GPUs.  The unit of work is one pipeline call (a batch of images with its CFG twin); units are dealt
round-robin to ranks, every rank runs whole units with replicated weights, and the only collective is
the gather of finished latents (NCCL over NVLink on GPUs, gloo in the CPU tests).  Because a unit's
arithmetic does not depend on which rank runs it, the gathered result is identical for any world size.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def units_for_rank(n_units: int, rank: int, world_size: int) -> List[int]:
    """Unit j goes to rank j % world_size (rank r runs r, r+G, r+2G, ...)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_units, world_size))


def seeds_of_unit(unit: int, images_per_unit: int) -> List[int]:
    """BASELINE configs[3]: unit j holds seeds 8j .. 8j+7."""
    return list(range(unit * images_per_unit, (unit + 1) * images_per_unit))


def unit_noise(unit: int, images_per_unit: int, shape: Sequence[int]) -> torch.Tensor:
    """Per-image CPU generators so that a seed's noise never depends on batch or rank."""
    out = []
    for s in seeds_of_unit(unit, images_per_unit):
        g = torch.Generator().manual_seed(s)
        out.append(torch.randn([1, *shape], generator=g))
    return torch.cat(out)


def run_sharded(generate: Callable[[int], torch.Tensor], n_units: int) -> torch.Tensor:
    """Run `generate(unit) -> [images_per_unit, ...]` for this rank's units and gather every unit's
    result on every rank, ordered by unit index.  Works without an initialised process group (1 rank)."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    if world > 1 and n_units < world:  # same verdict on every rank, before any work or collective: nobody hangs
        raise ValueError(f"every rank needs at least one unit (n_units={n_units} < world_size={world})")
    mine = units_for_rank(n_units, rank, world)
    local = [generate(u) for u in mine]
    if world == 1:
        return torch.stack(local) if local else torch.empty(0)
    per_rank = (n_units + world - 1) // world
    proto = local[0]
    buf = proto.new_zeros((per_rank, *proto.shape))
    for i, t in enumerate(local):
        buf[i] = t
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)  # the one collective of the path
    out = proto.new_zeros((n_units, *proto.shape))
    for r in range(world):
        for i, u in enumerate(units_for_rank(n_units, r, world)):
            out[u] = gathered[r][i]
    return out
