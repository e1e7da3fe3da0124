"""How the hot path spreads over the GPUs of a box: work units are independent, so ranks only split
an index range -- there is no data-path collective. One process per GPU, ``RANK`` / ``WORLD_SIZE`` from the
launcher (torchrun) or explicit arguments."""
from __future__ import annotations

import contextlib
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


def local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))


@contextlib.contextmanager
def process_group(need_collectives: bool = True):
    """What every multi-rank entry point runs under: bind this process to ITS GPU (``LOCAL_RANK``) and, when the
    launcher started several ranks (``WORLD_SIZE`` > 1) and the run exchanges anything (gradients, metric rows),
    initialise ``torch.distributed`` -- NCCL on a GPU box, gloo otherwise -- and tear it down again if it was created
    here. Rendezvous comes from the launcher's ``MASTER_ADDR`` / ``MASTER_PORT``. A group the caller already
    initialised (bench.py, tests) is used as it is and left alone."""
    import torch
    import torch.distributed as dist

    _, world = rank_world()
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank() % torch.cuda.device_count())
    created = False
    if world > 1 and need_collectives and dist.is_available() and not dist.is_initialized():
        if torch.cuda.is_available():
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        else:
            dist.init_process_group("gloo")
        created = True
    try:
        yield
    finally:
        if created:
            dist.destroy_process_group()


def require_process_group(what: str) -> None:
    """Ranks that shard a run by RANK / WORLD_SIZE but cannot talk to each other would each silently produce a partial
    result: refuse instead."""
    import torch.distributed as dist

    _, world = rank_world()
    if world > 1 and not (dist.is_available() and dist.is_initialized() and dist.get_world_size() == world):
        raise RuntimeError(f"{what}: WORLD_SIZE={world} but torch.distributed is not initialised with that many ranks; "
                           "run through the entry point (python -m cryovit.training...) or under shard.process_group()")
