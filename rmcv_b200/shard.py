"""Frame sharding across the GPUs of one box.

Frames are independent units (the reference's process_function keeps no cross-frame state on the hot path,
executable/main.cpp:167-176), so a batch is cut into contiguous per-rank slices and there is NO data-path collective.
torch.distributed is used only as plumbing: a barrier around the timed region and a MAX-reduce of the per-rank
device time / SUM-reduce of the per-rank unit counts.
"""
from __future__ import annotations

from typing import List, Tuple


def frame_slice(total_frames: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch owned by `rank`; sizes differ by at most one frame."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(total_frames), world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_slices(total_frames: int, world_size: int) -> List[Tuple[int, int]]:
    return [frame_slice(total_frames, world_size, r) for r in range(world_size)]


def reduce_timing(local_ms: float, local_units: int, dist=None, device=None) -> Tuple[float, int]:
    """(max over ranks of local_ms, sum over ranks of local_units).  `dist` = torch.distributed or None (single rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms), int(local_units)
    import torch
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([int(local_units)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), int(u.item())


def gather_counts(local_counts: List[int], dist=None, device=None) -> List[List[int]]:
    """Host-side concatenation of per-rank per-frame counts in rank order (the only cross-GPU step of the path)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(local_counts)]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(local_counts))
    return out
