"""``cryovit.models.CryoVIT`` for the hot path (models/cryovit.py:10-49 on top of models/base_model.py:20-112): the
Hydra target of ``configs/model/cryovit.yaml`` with the reference's constructor keywords, parameter names and call
surface, executing on the sm_100a head (``cryovit_b200.head.CryoVITHeadB200``)."""
from __future__ import annotations

import math
from typing import Any, Callable

import torch

from ..head import BLOCKS, CryoVITHeadB200, state_dict_keys
from .._lib import CryovitB200Error
from .config import instantiate


def default_state_dict(in_channels: int = 1536, seed: int | None = None) -> dict[str, torch.Tensor]:
    """Random parameters with torch's default distributions under the reference's names."""
    g = torch.Generator()
    if seed is not None:
        g.manual_seed(seed)
    else:
        g.seed()
    sd: dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        fan_in = cin * k[0] * k[1] * k[2]
        b = 1.0 / math.sqrt(fan_in)
        sd[name + ".weight"] = (torch.rand(cout, cin, *k, generator=g) * 2 - 1) * b
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * b

    conv("layers.0", 1024, in_channels, (1, 1, 1))
    for bi, (c1, c2, c3, _, _) in enumerate(BLOCKS):
        p = f"layers.{bi + 2}.layers."
        sd[p + "0.weight"], sd[p + "0.bias"] = torch.ones(c1), torch.zeros(c1)
        conv(p + "1", c2, c1, (3, 3, 3))
        conv(p + "3", c2, c2, (3, 3, 3))
        # ConvTranspose3d weight is [in, out, 1, 2, 2]; torch computes its fan_in from dim 1
        fan_in = c3 * 1 * 2 * 2
        b = 1.0 / math.sqrt(fan_in)
        sd[p + "5.weight"] = (torch.rand(c2, c3, 1, 2, 2, generator=g) * 2 - 1) * b
        sd[p + "5.bias"] = (torch.rand(c3, generator=g) * 2 - 1) * b
    conv("output_layer.0", 8, 8, (3, 3, 3))
    conv("output_layer.2", 1, 8, (3, 3, 3))
    assert set(sd) == set(state_dict_keys())
    return sd


class CryoVIT:
    """Drop-in for the reference's ``CryoVIT`` at inference / evaluation time.

    Constructor keywords are those of ``BaseModel`` (base_model.py:20-56): ``input_key, lr, weight_decay, losses,
    metrics, name, custom_kwargs``; ``in_channels`` generalises the hard-wired 1536 (BASELINE config 1).
    """

    def __init__(self, input_key: str = "dino_features", lr: float = 1e-4, weight_decay: float = 1e-3,
                 losses: dict[str, Any] | None = None, metrics: dict[str, Any] | None = None, name: str = "CryoVIT",
                 custom_kwargs: dict | None = None, in_channels: int = 1536, model_dir=None, seed: int | None = None, **_):
        self.name, self.input_key, self.lr, self.weight_decay = name, input_key, lr, weight_decay
        for k, v in (custom_kwargs or {}).items():
            setattr(self, k, v)
        self.loss_fns: dict[str, Callable] = {k: self._build(v) for k, v in (losses or {}).items()}
        self.metric_fns: dict[str, Any] = {k: self._build(v) for k, v in (metrics or {}).items()}
        self.in_channels = in_channels
        self._head = CryoVITHeadB200(in_channels).load_state_dict(default_state_dict(in_channels, seed))

    @staticmethod
    def _build(node):
        return instantiate(node) if isinstance(node, dict) and "_target_" in node else node

    # ------------------------------------------------------------------ nn.Module-like surface
    def state_dict(self) -> dict[str, torch.Tensor]:
        return self._head.state_dict()

    def load_state_dict(self, sd: dict[str, torch.Tensor], strict: bool = True):
        self._head.load_state_dict(sd, strict=strict)
        return self

    def cuda(self, device=None):
        self._head.cuda(device)
        return self

    def to(self, device):
        d = torch.device(device)
        if d.type != "cuda":
            raise CryovitB200Error("CryoVIT (B200) only runs on CUDA devices")
        return self.cuda(d.index)

    def eval(self):
        return self

    def forward_volume(self, x: torch.Tensor) -> torch.Tensor:
        return self._head.forward_volume(x)

    def forward(self, batch) -> torch.Tensor:
        return self._head.forward(batch)

    __call__ = forward

    # ------------------------------------------------------------------ evaluation (base_model.py:91-164,176-241)
    def _masked_predict(self, batch, use_mito_mask: bool = False) -> dict[str, torch.Tensor]:
        """Probabilities and the labels they are scored against. With ``use_mito_mask`` (reference
        base_model.py:91-111: granule experiments, batch size 1) voxels outside ``aux_data["labels/mito"] > 0`` are
        excluded too: their label becomes -1, which every fused reduction pass already ignores."""
        probs = self(batch)
        labels = batch.labels.to(probs.device)
        if use_mito_mask:
            aux = getattr(batch, "aux_data", None)
            if not aux or "labels/mito" not in aux:
                raise CryovitB200Error("Batch aux_data must contain 'labels/mito' key for mito masking.")
            mito = torch.as_tensor(aux["labels/mito"][0]).to(labels.device) > 0
            labels = torch.where(mito.view_as(labels[0]).expand_as(labels), labels, torch.full_like(labels, -1.0))
        return {"preds_full": probs, "labels": labels}

    def test_step(self, batch) -> dict[str, float]:
        """Losses and metrics over the voxels with label > -1 (one fused reduction pass per quantity); the reference's
        opt-in mito mask (base_model.py:192-199) applies when 'labels/mito' travels in aux_data."""
        aux = getattr(batch, "aux_data", None)
        use_mito = bool(aux) and "labels/mito" in aux and aux["labels/mito"] is not None and len(aux["labels/mito"]) > 0
        out = self._masked_predict(batch, use_mito_mask=use_mito)
        res = {}
        for k, fn in self.loss_fns.items():
            res[k] = float(fn(out["preds_full"], out["labels"]))
        for k, m in self.metric_fns.items():
            m.reset()
            m.update(out["preds_full"], out["labels"])
            res[k] = float(m.compute())
        return res
