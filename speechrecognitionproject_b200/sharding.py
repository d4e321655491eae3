"""Clip-level sharding for multi-GPU runs: clips are independent, so a
corpus is split into contiguous per-rank slices and no collective touches the data path."""
from __future__ import annotations


def shard_range(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of rank ``rank``; sizes differ by at most one clip."""
    if not 0 <= rank < world:
        raise ValueError("need 0 <= rank < world")
    base, rem = divmod(n_clips, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
