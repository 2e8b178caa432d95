"""Slice sharding for multi-GPU inference.

A slice (one [E, T, Y, X] image series with its own maps / mask) never interacts with another one in the
data-consistency path, so N GPUs = N processes over disjoint contiguous slice ranges and **no data-path
collective** (the reference does the same with nn.DataParallel scatter/gather, scripts/reconstruct.py:199-206).
torch.distributed is used only for the timing barrier / max-over-ranks and, optionally, to gather results.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def slice_range(rank: int, world: int, n_slices: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of the slices rank `rank` owns; sizes differ by at most one, order preserved."""
    if world <= 0 or not (0 <= rank < world) or n_slices < 0:
        raise ValueError(f"bad sharding request rank={rank} world={world} n_slices={n_slices}")
    base, extra = divmod(n_slices, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Max of a per-rank scalar (device-timed milliseconds) over the default process group; identity without one."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_slices(local: torch.Tensor, n_slices: int) -> Optional[torch.Tensor]:
    """Reassemble the per-rank results (dim 0 = slices) on rank 0 in slice order; other ranks get None.
    Host-side convenience for writing results — not part of the timed path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    if torch.is_complex(local):                            # collectives move real views (gloo has no complex gather)
        real = gather_slices(torch.view_as_real(local.contiguous()), n_slices)
        return None if real is None else torch.view_as_complex(real)
    sizes = [slice_range(r, world, n_slices) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bucket = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bucket, dst=0)
    if rank != 0:
        return None
    return torch.cat([bucket[r][: e - b] for r, (b, e) in enumerate(sizes)], dim=0)
