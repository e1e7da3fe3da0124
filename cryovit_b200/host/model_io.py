"""``.model`` files (reference utils.py:325-378 ``SavedModel`` / ``save_model``, :431-468 ``load_model``): a pickle of a
dataclass holding the model's name, type, label key, Hydra config node and state dict -- the format in which trained
heads are handed to ``cryovit infer`` (run/infer_model.py) and fine-tuned from (run/train_model.py ``ckpt_path``).

Files written by the reference pickle ``model_cfg`` as an ``omegaconf.DictConfig``. With omegaconf installed they
unpickle natively; without it (this image) the unpickler substitutes stand-ins for the omegaconf classes and the config is
read out of their pickled state (``_content`` / ``_val``), which is all this path needs: the ``_target_`` and the
constructor keywords. Files written here store ``model_cfg`` as the plain nested dict of ``cryovit_b200.host.config``.
Only ``cryovit.models.CryoVIT`` heads load; other targets raise (they are outside the B200 hot path)."""
from __future__ import annotations

import enum
import importlib
import pickle
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import torch

from .._lib import CryovitB200Error


class ModelType(enum.Enum):
    """types.py:49-55."""

    CRYOVIT = "cryovit"
    UNET3D = "unet3d"
    SAM2 = "sam2"
    MEDSAM = "medsam"


@dataclass
class SavedModel:
    """utils.py:325-351."""

    name: str
    model_type: ModelType
    label_key: str
    model_cfg: Any
    weights: dict[str, Any]


# pickled under the reference's class paths (cryovit/utils.py and cryovit/types.py re-export these objects), so that a
# file written here names the same classes as one written by the reference
SavedModel.__module__ = "cryovit.utils"
ModelType.__module__ = "cryovit.types"


class _Stub:
    """Stand-in for a class that cannot be imported while unpickling (omegaconf nodes): keeps the pickled state."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


def plain_config(node) -> Any:
    """Nested dict / list / scalar view of a config node: our own ``Cfg`` dicts, real omegaconf containers, or the
    stand-ins above (DictConfig / ListConfig keep their children in ``_content``, value nodes their value in ``_val``)."""
    try:
        from omegaconf import OmegaConf  # type: ignore

        if OmegaConf.is_config(node):
            return OmegaConf.to_container(node, resolve=True)
    except ImportError:
        pass
    if isinstance(node, dict):
        return {k: plain_config(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [plain_config(v) for v in node]
    d = getattr(node, "__dict__", None)
    if isinstance(d, dict):
        if "_content" in d:
            return plain_config(d["_content"])
        if "_val" in d:
            return plain_config(d["_val"])
    if isinstance(node, enum.Enum):
        return node.value
    return node


class _Unpickler(pickle.Unpickler):
    _OURS = {("cryovit.utils", "SavedModel"): SavedModel, ("cryovit.types", "ModelType"): ModelType}

    def find_class(self, module, name):
        if (module, name) in self._OURS:
            return self._OURS[(module, name)]
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            if module.split(".")[0] in ("omegaconf", "hydra", "cryovit"):
                return type(name, (_Stub,), {"__module__": module})
            raise


def save_model(model_name: str, label_key: str, model, model_cfg, save_path: Path | str) -> None:
    """utils.py:354-378. ``model_cfg`` is stored as a plain nested dict (readable with or without omegaconf)."""
    cfg = plain_config(model_cfg)
    stored: Any = cfg
    try:  # the reference reads ``model_cfg._target_``: keep it a DictConfig whenever omegaconf is there to make one
        from omegaconf import OmegaConf  # type: ignore

        stored = OmegaConf.create(cfg)
    except ImportError:
        pass
    saved = SavedModel(name=model_name, model_type=ModelType(str(cfg.get("name", "CryoVIT")).lower()), label_key=label_key,
                       model_cfg=stored, weights={k: v.detach().cpu() for k, v in model.state_dict().items()})
    Path(save_path).parent.mkdir(parents=True, exist_ok=True)
    with open(save_path, "wb") as fh:
        pickle.dump(saved, fh)


def load_model(model_path: Path | str, load_model: bool = True):
    """utils.py:431-468: ``(model | None, model_type, name, label_key)``. The model is the B200 ``cryovit.models.CryoVIT``
    with the file's weights loaded (not yet moved to a device)."""
    model_path = Path(model_path)
    if not model_path.exists():
        raise FileNotFoundError(f"Model file {model_path} does not exist.")
    with open(model_path, "rb") as fh:
        saved = _Unpickler(fh).load()
    model_type = saved.model_type if isinstance(saved.model_type, ModelType) else ModelType(plain_config(saved.model_type))
    model = None
    if load_model:
        cfg = plain_config(saved.model_cfg)
        target = cfg.get("_target_", "") if isinstance(cfg, dict) else ""
        if target not in ("cryovit.models.CryoVIT", "cryovit.models.cryovit.CryoVIT"):
            raise CryovitB200Error(f"{model_path}: model target {target!r} is outside the B200 hot path (CryoVIT head only)")
        weights = {k: torch.as_tensor(v) for k, v in saved.weights.items()}
        kwargs = {k: v for k, v in cfg.items() if k not in ("_target_", "_partial_")}
        kwargs.setdefault("in_channels", int(weights["layers.0.weight"].shape[1]))
        mod, _, attr = "cryovit.models.CryoVIT".rpartition(".")
        model = getattr(importlib.import_module(mod), attr)(**kwargs)
        model.load_state_dict(weights)
    return model, model_type, saved.name, saved.label_key
