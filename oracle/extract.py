"""fp32 restatement of the reference feature-extraction loop (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows src/cryovit/run/dino_features.py:31-64 (_dino_features): per batch of <= batch_size slices,
forward_features -> x_norm_patchtokens [B, Np, C] -> reshape (B, w, h, C) -> permute (C, B, w, h) ->
float16 -> concatenate along the slice axis. ``.cuda()`` is dropped (CPU oracle).
"""
from __future__ import annotations

import numpy as np
import torch


@torch.inference_mode()
def dino_features(data: torch.Tensor, model, batch_size: int) -> np.ndarray:
    w, h = np.array(data.shape[-2:]) // 14
    all_features = []
    for i in range(0, len(data), batch_size):
        if i + batch_size > len(data):
            batch_size = len(data) - i
        vec = data[i:i + batch_size]
        features = model.forward_features(vec)["x_norm_patchtokens"]
        features = features.reshape(features.shape[0], w, h, -1)
        features = features.permute([3, 0, 1, 2]).contiguous()
        all_features.append(features.to("cpu").half().numpy())
    return np.concatenate(all_features, axis=1)
