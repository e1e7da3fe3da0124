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

from . import nvtx, ops
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
    if D == 0:  # the reference concatenates an empty list of batches here (run/dino_features.py:64)
        raise ValueError("need at least one array to concatenate")
    if data.shape[-2] % 14 or data.shape[-1] % 14:
        raise CryovitB200Error(f"slice size {tuple(data.shape[-2:])} is not a multiple of the 14-pixel patch")
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
        with nvtx.span(f"extract.slices[{i}:{min(i + batch_size, D)}]"):
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


class _Ticket:
    """Handle of one submitted tomogram: ``result()`` blocks until its features are on the host."""

    def __init__(self, event: torch.cuda.Event, host: torch.Tensor, shape: tuple[int, ...]):
        self._event, self._host, self._shape = event, host, shape

    def result(self) -> np.ndarray:
        self._event.synchronize()
        return self._host[: int(np.prod(self._shape))].view(self._shape).numpy()


class TomogramFeatureStream:
    """Pipelined feature extraction over a sequence of tomograms: the host->device copy of tomogram i+1 and the
    device->host copy of the features of tomogram i-1 run on their own CUDA streams under the ViT of tomogram i,
    so PCIe time disappears from the steady state (the reference moves every batch synchronously on one stream,
    run/dino_features.py:47-62). Device and pinned host buffers are recycled over ``depth`` slots: the array a
    ticket returns stays valid until ``depth`` further tomograms have been submitted.
    """

    def __init__(self, model: DinoVisionTransformerB200, batch_size: int = 128, depth: int = 2):
        if model.device is None:
            raise CryovitB200Error("TomogramFeatureStream needs a model on a CUDA device")
        self.model, self.batch_size, self.depth = model, batch_size, max(2, depth)
        self._h2d, self._d2h = torch.cuda.Stream(model.device), torch.cuda.Stream(model.device)
        self._slots: list[dict] = [{} for _ in range(self.depth)]
        self._n = 0

    def _buffers(self, slot: dict, tomo_shape, dtype, feat_shape) -> None:
        n_in, n_out = int(np.prod(tomo_shape)), int(np.prod(feat_shape))
        dev = self.model.device
        if slot.get("in_dtype") != dtype or slot.get("n_in", 0) < n_in:
            slot["in_host"] = torch.empty(n_in, dtype=dtype, pin_memory=True)
            slot["in_dev"] = torch.empty(n_in, dtype=dtype, device=dev)
            slot["in_dtype"], slot["n_in"] = dtype, n_in
        if slot.get("n_out", 0) < n_out:
            slot["out_dev"] = torch.empty(n_out, dtype=torch.float16, device=dev)
            slot["out_host"] = torch.empty(n_out, dtype=torch.float16, pin_memory=True)
            slot["n_out"] = n_out

    @torch.inference_mode()
    def submit(self, tomo: np.ndarray | torch.Tensor) -> _Ticket:
        t = torch.from_numpy(tomo) if isinstance(tomo, np.ndarray) else tomo
        if t.dtype not in (torch.uint8, torch.float32):
            t = t.float()
        D, H, W = t.shape
        _, _, gh, gw = ops.patch_grid(H, W)
        feat_shape = (self.model.embed_dim, D, gh, gw)
        slot = self._slots[self._n % self.depth]
        self._n += 1
        if "done" in slot:
            slot["done"].synchronize()  # the slot's previous features have left the device (and its input is free)
        for sl in ([slot] if self._n > 1 else self._slots):  # first submit: size every slot (pinned allocation is slow)
            self._buffers(sl, (D, H, W), t.dtype, feat_shape)
        main = torch.cuda.current_stream(self.model.device)
        n_in = D * H * W
        if t.is_cuda:
            src_dev = t.contiguous()
        else:
            slot["in_host"][:n_in].copy_(t.reshape(-1))  # pageable -> pinned staging (host memcpy)
            with torch.cuda.stream(self._h2d):
                slot["in_dev"][:n_in].copy_(slot["in_host"][:n_in], non_blocking=True)
            main.wait_stream(self._h2d)
            src_dev = slot["in_dev"][:n_in].view(D, H, W)
        feats = slot["out_dev"][: int(np.prod(feat_shape))].view(feat_shape)
        extract_tomogram_device(src_dev, self.model, self.batch_size, out=feats)
        computed = torch.cuda.Event()
        computed.record(main)
        done = torch.cuda.Event()
        with torch.cuda.stream(self._d2h):
            self._d2h.wait_event(computed)
            slot["out_host"][: feats.numel()].copy_(feats.reshape(-1), non_blocking=True)
            done.record(self._d2h)
        slot["done"] = done
        return _Ticket(done, slot["out_host"], feat_shape)
