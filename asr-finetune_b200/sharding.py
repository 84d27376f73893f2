"""Per-rank clip sharding that mirrors the reference's DDP layout.

The reference shards by `ray.train.get_dataset_shard("train")` under `ScalingConfig(num_workers=N)`
(ref:finetune/training/trainers/trainers.py:785-791, ref:finetune/training/train_hyper.py:319-322): every worker
gets an equal-size, disjoint subset and collates its own `per_device_train_batch_size` batches.  The frontend is
embarrassingly parallel over clips, so there is no collective on the data path — only this index arithmetic.
"""
from __future__ import annotations

import os
from typing import Iterator, Optional, Tuple


def world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun / Ray Train environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def rank_shard(n_clips: int, rank: Optional[int] = None, world_size: Optional[int] = None,
               drop_remainder: bool = True) -> range:
    """Contiguous, equal-size block of clip indices owned by `rank` (remainder dropped, like an equal split)."""
    r, _, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    per = n_clips // world_size
    if drop_remainder:
        return range(rank * per, (rank + 1) * per)
    extra = n_clips % world_size
    start = rank * per + min(rank, extra)
    return range(start, start + per + (1 if rank < extra else 0))


def shard_batches(n_clips: int, batch_size: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                  drop_last: bool = False) -> Iterator[range]:
    """Iterate this rank's shard in `per_device_train_batch_size` batches of clip indices."""
    shard = rank_shard(n_clips, rank, world_size)
    for start in range(shard.start, shard.stop, batch_size):
        stop = min(start + batch_size, shard.stop)
        if drop_last and stop - start < batch_size:
            return
        yield range(start, stop)
