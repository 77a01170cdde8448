"""Data-parallel image sharding of the pseudo-labelling path: images are independent, every rank owns a
teacher replica and ``IMS_PER_BATCH_TARGET // world_size`` images (reference daod/data/build.py:323-331).  No
data-path collective exists; NCCL is used only by the student's DDP and the optional AdaBN statistic all-reduce."""
from __future__ import annotations

from typing import Tuple


def images_per_rank(total_batch_size: int, world_size: int) -> int:
    """reference daod/data/build.py:323-331: the global batch must divide evenly."""
    assert total_batch_size > 0 and total_batch_size % world_size == 0, \
        "Total batch size ({}) must be divisible by the number of gpus ({}).".format(total_batch_size, world_size)
    return total_batch_size // world_size


def shard_range(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``num_items`` independent images for ``rank`` (sizes differ by at most one)."""
    assert 0 <= rank < world_size
    base, rem = divmod(num_items, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
