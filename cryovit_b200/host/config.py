"""Configuration for the hot path, with the reference's keys (config.py:17-156,205-231 and configs/*.yaml).

The reference composes its YAML tree with Hydra. Hydra / OmegaConf are used when importable; this image has
neither, so :func:`compose` implements the subset the hot path's configs need: ``defaults`` lists (group entries,
``_self_``, schema nodes and ``override hydra/...`` lines are accepted and skipped), ``${a.b}`` interpolation, ``???``
mandatory values, and command-line overrides ``a.b=value`` / ``+a.b=value`` / ``group=choice`` -- the forms
``slurm_scripts/dino_features_job.sh`` passes.
"""
from __future__ import annotations

import copy
import importlib
import logging
import re
import sys
from functools import partial
from pathlib import Path
from typing import Any

import yaml

DINO_PATCH_SIZE = 14  # config.py:17
MISSING = "???"
tomogram_exts: list[str] = [".hdf", ".mrc"]  # config.py:15

# types.py:14-44 (Sample): attribute name -> display name. ``samples`` is the list of directory names.
SAMPLES: dict[str, str] = {
    "BACHD": "BACHD", "BACHD_Microtubules": "BACHD Microtubules", "dN17_BACHD": "dN17 BACHD", "Q109": "Q109",
    "Q109_Microtubules": "Q109 Microtubules", "Q18": "Q18", "Q18_Microtubules": "Q18 Microtubules", "Q20": "Q20",
    "Q53": "Q53", "Q53_KD": "Q53 PIAS1", "Q66": "Q66", "Q66_GRFS1": "Q66 GRFS1", "Q66_KD": "Q66 PIAS1",
    "WT": "Wild Type", "WT_Microtubules": "Wild Type Microtubules", "cancer": "Cancer", "AD": "AD",
    "AD_Abeta": "AD Abeta", "Aged": "Aged", "Young": "Young", "RGC_CM": "RGC CM", "RGC_control": "RGC Control",
    "RGC_naPP": "RGC naPP", "RGC_PP": "RGC PP", "CZI_Algae": "Algae", "CZI_Campy_C": "Campy C",
    "CZI_Campy_CDel": "Campy C-Deletion", "CZI_Campy_F": "Campy F", "CZI_Fibroblast": "Mouse Fibroblast",
}
samples: list[str] = list(SAMPLES)

# schema defaults of the structured nodes the YAML tree merges with (config.py:113-156)
SCHEMA_DEFAULTS: dict[str, dict] = {
    "paths": {"model_dir": MISSING, "data_dir": MISSING, "exp_dir": MISSING, "results_dir": MISSING,
              "tomo_name": "tomograms", "feature_name": "dino_features", "dino_name": "DINOv2", "sam_name": "SAM2",
              "csv_name": "csv", "split_name": "splits.csv"},
    "dino_features": {"batch_size": 128, "model_dir": MISSING, "paths": MISSING, "model": None, "datamodule": MISSING,
                      "sample": MISSING, "export_features": False, "use_sam": False},
}


class Cfg(dict):
    """dict with attribute access (what the runner needs of a DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x):
    if isinstance(x, dict):
        return Cfg({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


_SCI = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)[eE][+-]?\d+$")


def _numbers(node):
    """PyYAML (YAML 1.1) reads ``1e-4`` as a string; OmegaConf reads it as a float. Follow OmegaConf."""
    if isinstance(node, dict):
        return {k: _numbers(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_numbers(v) for v in node]
    if isinstance(node, str) and _SCI.match(node):
        return float(node)
    return node


def _load_yaml(path: Path) -> dict:
    with open(path) as fh:
        return _numbers(yaml.safe_load(fh) or {})


class _DirSource:
    """A directory of Hydra YAML files (``config_dir=``): node ``<group>/<choice>`` is ``<dir>/<group>/<choice>.yaml``."""

    def __init__(self, root: Path):
        self.root = Path(root)

    def exists(self, rel: str) -> bool:
        return (self.root / (rel + ".yaml")).exists()

    def load(self, rel: str) -> dict:
        path = self.root / (rel + ".yaml")
        if not path.exists():
            raise FileNotFoundError(f"config file {path} not found")
        return _load_yaml(path)


class _TreeSource:
    """The built-in tree (``config_tree.TREE``): the same nodes as data."""

    def __init__(self):
        from .config_tree import TREE

        self.tree = TREE

    def exists(self, rel: str) -> bool:
        return rel in self.tree

    def load(self, rel: str) -> dict:
        if rel not in self.tree:
            raise FileNotFoundError(f"config node {rel!r} not found in the built-in tree")
        return _numbers(copy.deepcopy(self.tree[rel]))


def _compose_file(config_dir, rel: str, choices: dict[str, str]) -> dict:
    """One config node with its ``defaults`` list resolved (group paths are relative to the node's own group)."""
    body = config_dir.load(rel)
    defaults = body.pop("defaults", ["_self_"])
    group_dir = str(Path(rel).parent) if "/" in rel else ""
    out: dict = {}
    self_done = False
    for entry in defaults:
        if entry == "_self_":
            _merge(out, body)
            self_done = True
        elif isinstance(entry, str):
            cand = (group_dir + "/" if group_dir else "") + entry
            if config_dir.exists(cand):  # a sibling node of the same group
                _merge(out, _compose_file(config_dir, cand, choices))
            # else: a ConfigStore schema node (base_env, dino_features_config ...): defaults come from SCHEMA_DEFAULTS
        elif isinstance(entry, dict):
            for group, choice in entry.items():
                if str(group).startswith("override "):
                    continue  # hydra logging overrides
                optional = str(group).startswith("optional ")  # Hydra: skip silently when the choice does not exist
                if optional:
                    group = str(group)[len("optional "):].strip()
                full_group = (group_dir + "/" if group_dir else "") + group
                for ch in (choice if isinstance(choice, list) else [choice]):
                    ch = choices.get(full_group, ch)
                    if isinstance(ch, str) and ch.startswith("${") and ch.endswith("}"):  # e.g. ``${model}``: another group's choice
                        ch = choices.get(ch[2:-1], choices.get("=" + ch[2:-1], ch))
                    if optional and not (isinstance(ch, str) and config_dir.exists(f"{full_group}/{ch}")):
                        continue
                    if ch == MISSING:  # mandatory group (``model: ???``) not chosen: reported by the validator
                        out[group] = MISSING
                        continue
                    if not isinstance(choice, list):
                        choices.setdefault("=" + full_group, ch)  # recorded for ${hydra:runtime.choices.<group>}
                    sub = _compose_file(config_dir, f"{full_group}/{ch}", choices)
                    node = out.setdefault(group, {})
                    if isinstance(choice, list):  # list-valued defaults (losses, metrics): one sub-key per choice
                        node[ch] = sub.get(ch, sub) if isinstance(sub, dict) else sub
                    else:
                        _merge(node, sub)
    if not self_done:
        _merge(out, body)
    return out


_INTERP = re.compile(r"\$\{([^}]+)\}")


def _lookup(root: dict, dotted: str):
    cur: Any = root
    if dotted.startswith("hydra:"):  # the one resolver the reference's YAML uses: ${hydra:runtime.choices.<group>}
        dotted = "hydra." + dotted[len("hydra:"):]
    for part in dotted.split("."):
        if not isinstance(cur, dict) or part not in cur:
            return MISSING  # points into a mandatory value that was not given: stays "???" for the validator
        cur = cur[part]
    return cur


def _resolve(node, root, depth=0):
    if depth > 16:
        raise ValueError("interpolation cycle")
    if isinstance(node, dict):
        for k in list(node):
            node[k] = _resolve(node[k], root, depth)
        return node
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str) and "${" in node:
        m = _INTERP.fullmatch(node)
        if m:
            return _resolve(_lookup(root, m.group(1)), root, depth + 1)
        return _resolve(_INTERP.sub(lambda mm: str(_resolve(_lookup(root, mm.group(1)), root, depth + 1)), node), root, depth + 1)
    return node


def _parse_value(text: str):
    if text in ("null", "None", "~"):
        return None
    try:
        return _numbers(yaml.safe_load(text))
    except yaml.YAMLError:
        return text


def compose(config_name: str, overrides: list[str] | None = None, config_dir: Path | None = None) -> Cfg:
    """Compose node ``config_name`` of the built-in tree (or ``<config_dir>/<config_name>.yaml`` of a directory of Hydra
    YAML files) with its defaults and apply overrides."""
    config_dir = _DirSource(config_dir) if config_dir else _TreeSource()
    overrides = list(overrides or [])
    choices: dict[str, str] = {}
    values: list[tuple[str, Any]] = []
    for ov in overrides:
        if "=" not in ov:
            raise ValueError(f"override {ov!r} is not of the form key=value")
        key, val = ov.split("=", 1)
        key = key.lstrip("+")
        if key == "experiments":
            values.append((key, val))
        elif config_dir.exists(f"{key}/{val}"):
            choices[key] = val  # config-group choice, e.g. paths=default
        else:
            values.append((key, _parse_value(val)))
    # ``+experiments=<name>``: a ``# @package _global_`` file whose body merges at the root and whose defaults may
    # ``override /<group>: <choice>`` (configs/experiments/*.yaml); command-line group choices still win
    experiment_bodies: list[dict] = []
    for key, val in list(values):
        if key == "experiments":
            body = config_dir.load(f"experiments/{val}")
            for entry in body.pop("defaults", []):
                if isinstance(entry, dict):
                    for g, ch in entry.items():
                        if str(g).startswith("override /"):
                            choices.setdefault(str(g)[len("override /"):], ch)
            body.pop("hydra", None)  # multirun sweeps are a launcher feature, not part of the composed config
            experiment_bodies.append(body)
            values.remove((key, val))
    cfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in SCHEMA_DEFAULTS.get(config_name, {}).items()}
    _merge(cfg, _compose_file(config_dir, config_name, choices))
    for body in experiment_bodies:
        _merge(cfg, body)
    cfg["hydra"] = {"runtime": {"choices": {k[1:]: v for k, v in choices.items() if k.startswith("=")}}}
    if isinstance(cfg.get("paths"), dict):
        cfg["paths"] = _merge(dict(SCHEMA_DEFAULTS["paths"]), cfg["paths"])
    for key, val in values:
        cur = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            if not isinstance(cur.get(p), dict):
                cur[p] = {}
            cur = cur[p]
        cur[parts[-1]] = val
    cfg = _resolve(cfg, cfg)
    cfg.pop("hydra", None)
    return _wrap(cfg)


def missing_keys(cfg: dict, prefix: str = "") -> list[str]:
    out = []
    for k, v in cfg.items():
        if isinstance(v, dict):
            out += missing_keys(v, f"{prefix}{k}.")
        elif v == MISSING:
            out.append(prefix + k)
    return out


def validate_dino_config(cfg: dict) -> None:
    """config.py:205-231: list the mandatory parameters that are still ``???`` and exit(1)."""
    miss = missing_keys(cfg)
    if miss:
        msg = ["The following parameters were missing from dino_features.yaml"]
        msg += [f"{i}. {k}" for i, k in enumerate(miss, 1)]
        logging.error("\n".join(msg))
        sys.exit(1)


def validate_experiment_config(cfg: dict, config_name: str) -> None:
    """config.py:205-231 for train_model / eval_model: the same listing of ``???`` keys, then exit(1)."""
    miss = missing_keys(cfg)
    if miss:
        msg = [f"The following parameters were missing from {config_name}.yaml"]
        msg += [f"{i}. {k}" for i, k in enumerate(miss, 1)]
        logging.error("\n".join(msg))
        sys.exit(1)


def instantiate(node: dict, **kwargs):
    """hydra.utils.instantiate for the two forms the hot path uses: ``_target_`` (+ ``_partial_``)."""
    node = dict(node)
    target = node.pop("_target_")
    is_partial = bool(node.pop("_partial_", False))
    mod, _, attr = target.rpartition(".")
    fn = getattr(importlib.import_module(mod), attr)
    node.update(kwargs)
    return partial(fn, **node) if is_partial else fn(**node)
