"""fp32 restatement of the reference slice pre-processing (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows src/cryovit/datasets/vit_dataset.py:
  _load_tomogram  :71-88   uint8 -> float32 / 255, float data unchanged
  _dino_transform :90-123  edge-pad H, W up to multiples of 16; add channel dim and repeat to 3 channels;
                           F.interpolate(scale_factor=(14/16, 14/16), mode="bicubic")
The ImageNet Normalize built in __init__ (:39) is never applied by the reference, so it is not applied here.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

DINO_PATCH_SIZE = 14  # src/cryovit/config.py:17


def load_tomogram(data: np.ndarray) -> np.ndarray:
    if data.dtype == np.uint8:
        data = data.astype(np.float32) / 255.0
    return data


def dino_transform(data: np.ndarray) -> torch.Tensor:
    """[D, H, W] float -> [D, 3, H', W'] float32."""
    scale = (DINO_PATCH_SIZE / 16, DINO_PATCH_SIZE / 16)
    _, h, w = data.shape
    H = int(np.ceil(h / 16) * 16)
    W = int(np.ceil(w / 16) * 16)
    if h != H or w != W:
        data = np.pad(data, ((0, 0), (0, H - h), (0, W - w)), mode="edge")
    x = np.repeat(np.expand_dims(data, axis=1), 3, axis=1)
    return F.interpolate(torch.from_numpy(x).float(), scale_factor=scale, mode="bicubic")
