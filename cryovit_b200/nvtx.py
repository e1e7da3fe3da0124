"""Optional NVTX ranges around the stages of the hot path (the reference has none; nsys / ncu
``--nvtx`` timelines of this package become readable with them). Off unless ``CRYOVIT_B200_NVTX=1``: the disabled
``span`` is a shared no-op context manager, so the ranges cost nothing in the measured path."""
from __future__ import annotations

import contextlib
import os

ENABLED = os.environ.get("CRYOVIT_B200_NVTX", "0") == "1"
_NULL = contextlib.nullcontext()


def span(name: str):
    if not ENABLED:
        return _NULL
    import torch

    return torch.cuda.nvtx.range(name)
