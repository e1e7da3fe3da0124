"""Tomogram in, segmentation out, without leaving the GPU.

The reference decouples its two halves through disk: ``dino_features`` writes a 403 MB fp16 feature volume per
tomogram (run/dino_features.py:148-153), ``infer_model`` / ``eval_model`` read it back, up-cast it to fp32 and run the
head (datasets/tomo_dataset.py:110-123, datamodules/utils.py:33-38, run/infer_model.py:18-85). For pure inference
the volume never needs to exist outside HBM: the ViT's write-out kernel produces exactly the (C, D, h, w) fp16 layout
the head's first kernel consumes. ``PredictionWriter`` semantics are kept: mask = probabilities >= threshold, uint8
(models/callbacks.py:61-109).
"""
from __future__ import annotations

import numpy as np
import torch

from . import extract
from ._lib import CryovitB200Error
from .head import CryoVITHeadB200
from .vit import DinoVisionTransformerB200


@torch.inference_mode()
def segment_tomogram_device(tomo: torch.Tensor, vit: DinoVisionTransformerB200, head: CryoVITHeadB200, batch_size: int = 128,
                            want_features: bool = False):
    """tomo: CUDA [D, H, W] uint8 | float32 in [0, 1]. Returns (probabilities fp32 [D, 16h, 16w], features or None);
    the probability volume covers the tomogram padded up to multiples of 16 (crop with [:, :H, :W])."""
    if tomo.dim() != 3 or not tomo.is_cuda:
        raise CryovitB200Error("segment_tomogram_device expects a CUDA [D, H, W] tensor")
    feats = extract.extract_tomogram_device(tomo.contiguous(), vit, batch_size)
    _, probs = head.segment_volume(feats, want_logits=False)
    return probs, (feats if want_features else None)


@torch.inference_mode()
def segment_tomogram(tomo: np.ndarray | torch.Tensor, vit: DinoVisionTransformerB200, head: CryoVITHeadB200,
                     batch_size: int = 128, threshold: float = 0.5) -> np.ndarray:
    """Host entry: (D, H, W) uint8 | float32 tomogram -> uint8 mask (D, H, W), ``probs >= threshold``."""
    t = torch.from_numpy(tomo) if isinstance(tomo, np.ndarray) else tomo
    if t.dtype not in (torch.uint8, torch.float32):
        t = t.float()
    D, H, W = t.shape
    dev = t.contiguous().pin_memory().to(vit.device, non_blocking=True) if not t.is_cuda else t
    probs, _ = segment_tomogram_device(dev, vit, head, batch_size)
    mask = (probs[:, :H, :W] >= threshold).to(torch.uint8)
    host = torch.empty(mask.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(mask, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()
