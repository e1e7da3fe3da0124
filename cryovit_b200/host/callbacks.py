"""Evaluation writers with the reference's file formats (models/callbacks.py:15-59 ``TestPredictionWriter``, :112-206
``CsvWriter``). They consume a ``BatchedModelResult`` (types.py) -- here the plain dataclass of ``eval_model``."""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import numpy as np
import pandas as pd

from . import hdf


@dataclass
class BatchedModelResult:
    """types.py BatchedModelResult: everything ``test_step`` returns for one batch (base_model.py:176-241)."""

    num_tomos: int
    samples: list[str]
    tomo_names: list[str]
    split_id: list[int] | None
    data: list[np.ndarray]
    label: list[np.ndarray]
    preds: list[np.ndarray]
    losses: dict[str, float] = field(default_factory=dict)
    metrics: dict[str, float] = field(default_factory=dict)
    aux_data: dict[str, Any] | None = None


class TestPredictionWriter:
    """``<results_dir>/<sample>/<tomo_name>`` with ``data`` (raw), ``<label_key>`` (gzip), ``<label_key>_preds``
    (float probabilities, gzip) -- callbacks.py:41-59."""

    __test__ = False  # not a pytest class

    def __init__(self, results_dir: Path | str, label_key: str, **_kwargs):
        self.results_dir, self.label_key = Path(results_dir), label_key

    def on_test_batch_end(self, outputs: BatchedModelResult) -> None:
        for n in range(outputs.num_tomos):
            out = self.results_dir / outputs.samples[n] / outputs.tomo_names[n]
            hdf.write_tomogram(out, {"data": outputs.data[n], self.label_key: outputs.label[n],
                                     f"{self.label_key}_preds": outputs.preds[n]}, uncompressed=("data",))


class CsvWriter:
    """One row per test tomogram in ``<results_dir>/<sample>[_<split_id>].csv``: ``sample, tomo_name, <metrics...>,
    [split_id]``; an existing row of the same tomogram is replaced (callbacks.py:141-206)."""

    def __init__(self, results_dir: Path | str, **_kwargs):
        self.results_dir = Path(results_dir)
        self.results_dir.mkdir(parents=True, exist_ok=True)

    def on_test_batch_end(self, outputs: BatchedModelResult) -> None:
        assert outputs.num_tomos == 1, "TestPredictionWriter only supports single-tomogram batches."
        sample, tomo_name = outputs.samples[0], outputs.tomo_names[0]
        split_id = outputs.split_id[0] if outputs.split_id is not None else None
        path = self.results_dir / f"{sample}{'' if split_id is None else f'_{split_id}'}.csv"
        columns = ["sample", "tomo_name"] + list(outputs.metrics) + (["split_id"] if split_id is not None else [])
        df = pd.read_csv(path) if path.exists() else pd.DataFrame(columns=columns)
        match = (df["tomo_name"] == tomo_name) & (df["sample"] == sample)
        if split_id is not None and "split_id" in df.columns:
            match = match & (df["split_id"] == split_id)
        if match.any():
            logging.warning("Data with sample %s, name %s, and split %s already has an entry. Replacing %d rows...",
                            sample, tomo_name, split_id, int(match.sum()))
            df = df[~match]
        row: dict[str, Any] = {"sample": sample, "tomo_name": tomo_name}
        row.update({k: [v] for k, v in outputs.metrics.items()})
        if split_id is not None:
            row["split_id"] = [split_id]
        row_df = pd.DataFrame(row)
        df = row_df if df.empty else pd.concat([df, row_df], ignore_index=True)
        df.to_csv(path, mode="w", index=False)
