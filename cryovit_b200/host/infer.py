"""Segmentation inference over tomogram files with the reference's output layout (run/infer_model.py:18-85 with
models/callbacks.py:61-109 ``PredictionWriter``): for every input file a ``<result_dir>/<tomo stem>.hdf`` holding
``data`` (float32, gzip) and ``<label_key>_preds`` (uint8 mask ``probabilities >= threshold``, gzip).

The reference needs the feature file to exist on disk first; here a file that already carries ``dino_features`` is
fed to the head directly, any other goes tomogram -> ViT -> head without the features ever leaving HBM
(``cryovit_b200.pipeline``). Files are dealt round-robin to the ranks of a torchrun launch; no collective."""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch

from ..head import CryoVITHeadB200
from ..pipeline import segment_tomogram_device
from ..vit import DinoVisionTransformerB200
from . import hdf
from .shard import shard_round_robin


@torch.inference_mode()
def run_inference(data_files: list[Path | str], head: CryoVITHeadB200, result_dir: Path | str, threshold: float = 0.5,
                  label_key: str = "mito", vit: DinoVisionTransformerB200 | None = None, batch_size: int = 128) -> list[Path]:
    """Returns the result paths this rank wrote (the reference returns ``pred_writer.result_paths``)."""
    result_dir = Path(result_dir)
    out_paths: list[Path] = []
    for f in shard_round_robin([Path(p) for p in data_files]):
        keys = hdf.list_keys(f)
        data = hdf.read_tomogram(f, keys=["data"])["data"]
        D, H, W = data.shape
        if "dino_features" in keys:
            feats = torch.from_numpy(hdf.read_tomogram(f, keys=["dino_features"])["dino_features"]).to(head.device)
            _, probs = head.segment_volume(feats, want_logits=False)
        else:
            if vit is None:
                raise ValueError(f"{f} has no dino_features and no ViT was given")
            t = torch.from_numpy(np.ascontiguousarray(data if data.dtype in (np.uint8, np.float32) else data.astype(np.float32)))
            probs, _ = segment_tomogram_device(t.pin_memory().to(vit.device, non_blocking=True), vit, head, batch_size)
        segs = (probs[:, :H, :W] >= threshold).to(torch.uint8).cpu().numpy()
        as_f32 = data.astype(np.float32) / 255.0 if data.dtype == np.uint8 else data.astype(np.float32)
        path = (result_dir / f.name).with_suffix(".hdf")
        hdf.write_tomogram(path, {"data": as_f32, f"{label_key}_preds": segs}, uncompressed=())
        out_paths.append(path)
        logging.info("wrote %s (%d positive voxels)", path, int(segs.sum()))
    return out_paths
