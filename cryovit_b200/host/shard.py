"""How the hot path spreads over the GPUs of a box: work units are independent, so ranks only split
an index range -- there is no data-path collective. One process per GPU, ``RANK`` / ``WORLD_SIZE`` from the
launcher (torchrun) or explicit arguments."""
from __future__ import annotations

import os


def rank_world(rank: int | None = None, world: int | None = None) -> tuple[int, int]:
    r = int(os.environ.get("RANK", "0")) if rank is None else rank
    w = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    if not 0 <= r < w:
        raise ValueError(f"rank {r} outside world of {w}")
    return r, w


def shard_range(n: int, rank: int | None = None, world: int | None = None) -> range:
    """Contiguous, balanced slice range of rank: sizes differ by at most one, earlier ranks take the extra unit.
    Used for the z-slices of ONE tomogram: rank r writes features[:, d0:d1] and nothing else."""
    r, w = rank_world(rank, world)
    base, extra = divmod(n, w)
    start = r * base + min(r, extra)
    return range(start, start + base + (1 if r < extra else 0))


def shard_round_robin(items: list, rank: int | None = None, world: int | None = None) -> list:
    """Whole tomograms (feature extraction over a dataset, head inference): item i goes to rank i mod world."""
    r, w = rank_world(rank, world)
    return list(items[r::w])
