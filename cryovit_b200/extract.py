"""Feature extraction over a tomogram: the host side of reference run/dino_features.py:31-64.

Two entry points:
  * ``_dino_features(data, model, batch_size)`` -- same name, arguments and result as the reference function
    (seam B3): pre-processed float slices [D, 3, H', W'] in, ``np.float16 (C, D, H'/14, W'/14)`` out.
  * ``extract_tomogram(tomo, model, batch_size)`` -- the fused path from the RAW tomogram ([D, H, W] uint8 or
    float in [0, 1], what VITDataset._load_tomogram reads from the HDF ``data`` key): pre-processing runs on the
    GPU too, so only 1 byte per voxel crosses PCIe instead of 12 pre-processed floats.
Both keep the whole (C, D, h, w) fp16 volume on the device and make one pinned device->host copy at the end
(the reference makes one pageable fp32 copy + CPU cast per batch).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import CryovitB200Error
from .vit import DinoVisionTransformerB200


def _features_buffer(model: DinoVisionTransformerB200, D: int, gh: int, gw: int) -> torch.Tensor:
    return torch.empty(model.embed_dim, D, gh, gw, device=model.device, dtype=torch.float16)


def _to_host(features: torch.Tensor) -> np.ndarray:
    host = torch.empty(features.shape, dtype=torch.float16, pin_memory=True)
    host.copy_(features, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


@torch.inference_mode()
def _dino_features(data: torch.Tensor, model: DinoVisionTransformerB200, batch_size: int) -> np.ndarray:
    if not isinstance(model, DinoVisionTransformerB200):
        raise CryovitB200Error("_dino_features needs a DinoVisionTransformerB200 (build_model(...).cuda())")
    if data.dim() != 4 or data.shape[1] != 3:
        raise CryovitB200Error(f"expected [D, 3, H, W], got {tuple(data.shape)}")
    D = data.shape[0]
    w, h = np.array(data.shape[-2:]) // 14  # the reference's (confusingly named) rows, cols
    feats = _features_buffer(model, D, int(w), int(h))
    src = data if data.is_cuda else data.pin_memory()
    for i in range(0, D, batch_size):
        vec = src[i:i + batch_size].to(model.device, torch.float32, non_blocking=True)
        model.extract_preprocessed_into(vec, feats, i)
    return _to_host(feats)


@torch.inference_mode()
def extract_tomogram_device(tomo: torch.Tensor, model: DinoVisionTransformerB200, batch_size: int,
                            out: torch.Tensor | None = None) -> torch.Tensor:
    """tomo: CUDA tensor [D, H, W] uint8 / float32. Returns the fp16 (C, D, h, w) volume on the device."""
    D, H, W = tomo.shape
    _, _, gh, gw = ops.patch_grid(H, W)
    feats = out if out is not None else _features_buffer(model, D, gh, gw)
    for i in range(0, D, batch_size):
        model.extract_into(tomo[i:i + batch_size], feats, i)
    return feats


@torch.inference_mode()
def extract_tomogram(tomo: np.ndarray | torch.Tensor, model: DinoVisionTransformerB200, batch_size: int = 128) -> np.ndarray:
    t = torch.from_numpy(tomo) if isinstance(tomo, np.ndarray) else tomo
    if t.dtype not in (torch.uint8, torch.float32):
        t = t.float()
    if not t.is_cuda:
        t = t.contiguous().pin_memory().to(model.device, non_blocking=True)
    return _to_host(extract_tomogram_device(t.contiguous(), model, batch_size))
