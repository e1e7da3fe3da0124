"""Split-file data modules on either side of the head (reference datamodules/base_datamodule.py:14-128,
single_sample_datamodule.py:8-105, multi_sample_datamodule.py:8-103): which tomograms of ``splits.csv`` are the train /
validation / test / predict sets of an experiment, and the loaders over them.

The selection rules are the reference's, row for row. The loaders are plain iterables instead of
``torch.utils.data.DataLoader`` workers: an item is one whole tomogram (``batch_size`` 1, dataloader/default.yaml:7),
read by a background thread one item ahead of the GPU, collated by the reference's ``collate_fn`` and -- under
``torchrun`` -- dealt round-robin to the ranks from a common (seeded) order so that every rank takes the same number
of steps (the gradient all-reduce of head training is collective)."""
from __future__ import annotations

import threading
from collections.abc import Callable, Iterator
from pathlib import Path
from queue import Queue

import numpy as np
import pandas as pd

from .datasets import BatchedTomogramData, collate_fn
from .shard import rank_world


class TomoLoader:
    """``for batch in loader``: ``collate_fn([dataset[i]])`` for this rank's share of the (optionally shuffled) items.

    ``shuffle=True`` draws a fresh permutation per epoch from ``np.random.default_rng(seed + epoch)``: identical on
    every rank, so the round-robin shares are disjoint. ``drop_uneven`` (training) truncates the order to a multiple
    of the world size; evaluation keeps every item (ranks may differ by one item, there is no collective there)."""

    def __init__(self, dataset, shuffle: bool = False, collate: Callable = collate_fn, seed: int = 42, drop_uneven: bool | None = None,
                 prefetch: int = 1, **_ignored):
        self.dataset, self.shuffle, self.collate, self.seed = dataset, shuffle, collate, seed
        self.drop_uneven = shuffle if drop_uneven is None else drop_uneven
        self.prefetch = prefetch
        self.epoch = 0

    def indices(self) -> list[int]:
        rank, world = rank_world()
        n = len(self.dataset)
        order = np.random.default_rng(self.seed + self.epoch).permutation(n) if self.shuffle else np.arange(n)
        if self.drop_uneven:
            order = order[: n // world * world]
        return [int(i) for i in order[rank::world]]

    def __len__(self) -> int:
        return len(self.indices())

    def __iter__(self) -> Iterator[BatchedTomogramData]:
        idx = self.indices()
        self.epoch += 1
        if self.prefetch <= 0:
            for i in idx:
                yield self.collate([self.dataset[i]])
            return
        q: Queue = Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def work():
            try:
                for i in idx:
                    if stop.is_set():
                        return
                    q.put(("ok", self.collate([self.dataset[i]])))
                q.put(("end", None))
            except BaseException as e:  # noqa: BLE001  (re-raised in the consumer)
                q.put(("err", e))

        t = threading.Thread(target=work, daemon=True)
        t.start()
        try:
            while True:
                kind, val = q.get()
                if kind == "end":
                    return
                if kind == "err":
                    raise val
                yield val
        finally:
            stop.set()


class BaseDataModule:
    """base_datamodule.py:14-128. ``dataset_fn(records, train=...)`` builds a dataset from a dataframe of records,
    ``dataloader_fn(dataset, shuffle=..., collate_fn=...)`` a loader (default: :class:`TomoLoader`)."""

    def __init__(self, split_file: Path | str, dataset_fn: Callable, dataloader_fn: Callable | None = None, **_kwargs):
        self.dataset_fn = dataset_fn
        self.dataloader_fn = dataloader_fn or (lambda ds, shuffle, collate_fn: TomoLoader(ds, shuffle=shuffle, collate=collate_fn))
        self.split_file = Path(split_file)
        self.record_df = pd.read_csv(self.split_file)

    def train_df(self) -> pd.DataFrame:
        raise NotImplementedError

    def val_df(self) -> pd.DataFrame:
        raise NotImplementedError

    def test_df(self) -> pd.DataFrame:
        raise NotImplementedError

    def predict_df(self) -> pd.DataFrame:
        raise NotImplementedError

    def _loader(self, records: pd.DataFrame, what: str, train: bool):
        if records.empty:
            raise ValueError(f"No {what} data found in the provided split file.")
        dataset = self.dataset_fn(records, train=train)
        return self.dataloader_fn(dataset, shuffle=train, collate_fn=collate_fn)

    def train_dataloader(self):
        return self._loader(self.train_df(), "training", True)

    def val_dataloader(self):
        return self._loader(self.val_df(), "validation", False)

    def test_dataloader(self):
        return self._loader(self.test_df(), "testing", False)

    def predict_dataloader(self):
        return self._loader(self.predict_df(), "prediction", False)


class SingleSampleDataModule(BaseDataModule):
    """single_sample_datamodule.py:8-105: train on one sample (all splits but ``split_id``), validate on ``split_id``,
    test on the validation set or on the whole of another sample."""

    def __init__(self, sample: list[str], split_id: int | None, split_key: str, test_sample: list[str] | None = None, **kwargs):
        super().__init__(**kwargs)
        assert len(sample) == 1, f"Single sample 'sample' should be a single string list. Got {sample} instead."
        assert test_sample is None or len(test_sample) == 1, (
            f"Single sample 'test_sample' should be a single string list or None. Got {test_sample} instead.")
        self.sample, self.split_id, self.split_key = sample[0], split_id, split_key
        self.test_sample = test_sample[0] if test_sample is not None else None

    def _mine(self) -> pd.Series:
        return self.record_df["sample"] == self.sample

    def train_df(self) -> pd.DataFrame:
        if self.split_id is not None:
            return self.record_df[(self.record_df[self.split_key] != self.split_id) & self._mine()]
        return self.record_df[self._mine()][["sample", "tomo_name"]]

    def val_df(self) -> pd.DataFrame:
        if self.split_id is None:  # validate on the train set
            return self.train_df()
        return self.record_df[(self.record_df[self.split_key] == self.split_id) & self._mine()]

    def test_df(self) -> pd.DataFrame:
        if self.test_sample is None:
            return self.val_df()
        return self.record_df[self.record_df["sample"] == self.test_sample][["sample", "tomo_name"]]

    def predict_df(self) -> pd.DataFrame:
        return self.record_df[self._mine()][["sample", "tomo_name"]]


class MultiSampleDataModule(SingleSampleDataModule):
    """multi_sample_datamodule.py:8-103: the same rules over a LIST of training samples (BASELINE config 5:
    ``+experiments=multi_*``) and an optional list of test samples."""

    def __init__(self, sample: list[str], split_id: int | None, split_key: str | None, test_sample: list[str] | None = None, **kwargs):
        BaseDataModule.__init__(self, **kwargs)
        assert isinstance(sample, list), f"Multi sample 'sample' should be a list. Got {sample} instead."
        assert test_sample is None or isinstance(test_sample, list), (
            f"Multi sample 'test_sample' should be None or a list. Got {test_sample} instead.")
        self.sample, self.split_id, self.split_key, self.test_sample = sample, split_id, split_key, test_sample

    def _mine(self) -> pd.Series:
        return self.record_df["sample"].isin(self.sample)

    def test_df(self) -> pd.DataFrame:
        if self.test_sample is None:
            return self.val_df()
        return self.record_df[self.record_df["sample"].isin(self.test_sample)][["sample", "tomo_name"]]
