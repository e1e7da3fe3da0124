"""fp32 restatement of the masked prediction, loss and metrics (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows src/cryovit/models/base_model.py:91-112 (_masked_predict: mask = labels > -1, masked_select),
losses.py:17-32 (DiceLoss, eps 1e-3), metrics.py:30-53 (DiceMetric: pred < thr -> 0 else 1, eps 1e-3) and
metrics.py:69-93 (F1Metric: pred > 0.5, eps 1e-6), callbacks.py:100 (mask = preds >= threshold).
"""
from __future__ import annotations

import torch


def masked_select(y_pred_full: torch.Tensor, y_true: torch.Tensor):
    mask = y_true > -1.0
    return torch.masked_select(y_pred_full, mask).view(-1, 1), torch.masked_select(y_true, mask).view(-1, 1)


def dice_loss(y_pred, y_true):
    inter = torch.sum(y_true * y_pred)
    denom = torch.sum(y_true) + torch.sum(y_pred)
    return 1 - (2 * inter) / (denom + 1e-3)


def dice_metric(y_pred, y_true, threshold: float = 0.5):
    y_pred = torch.where(y_pred < threshold, 0.0, 1.0)
    inter = torch.sum(y_true * y_pred)
    denom = torch.sum(y_true) + torch.sum(y_pred)
    return 2 * inter / (denom + 1e-3)


def f1_metric(y_pred, y_true):
    y_pred = (y_pred > 0.5).float()
    tp = torch.sum(y_true * y_pred)
    fp = torch.sum((1 - y_true) * y_pred)
    fn = torch.sum(y_true * (1 - y_pred))
    precision = tp / (tp + fp + 1e-6)
    recall = tp / (tp + fn + 1e-6)
    return 2 * (precision * recall) / (precision + recall + 1e-6)


def prediction_mask(preds: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    return (preds >= threshold).to(torch.uint8)
