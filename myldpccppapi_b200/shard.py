"""Batch sharding of independent codewords over the GPUs of one box (no collective).

The reference decodes a stream codeword by codeword (MyLdpc.cpp:694) / chunk by chunk
(MyLdpc.cpp:577-616); codewords never interact, so GPU g simply decodes the contiguous range
shard_range(ncw, g, G) and its bytes land in the matching slice of the caller's buffer.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(ncw: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous [begin, end) of codewords for `rank`.  `align` keeps shard boundaries on a
    multiple of `align` codewords (use 8 when K is not a multiple of 8 so bytes stay whole)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    units = (ncw + align - 1) // align
    b = units * rank // world * align
    e = units * (rank + 1) // world * align
    return min(b, ncw), min(e, ncw)


def shard_ranges(ncw: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    return [shard_range(ncw, r, world, align) for r in range(world)]
