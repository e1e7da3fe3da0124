"""Feature-extraction runner with the reference's functions and file layout (run/dino_features.py:31-64,109-205,
304-350): ``_dino_features``, ``_save_data``, ``_process_sample``, ``run_trainer``.

Differences that are deliberate, all on the fast side of the same contract:
  * the model is ``cryovit_b200.vit.build_model("dinov2_vitg14_reg")`` (upstream checkpoint keys) instead of
    ``torch.hub.load`` (:336). Like the reference, which fails when torch.hub cannot produce the model (:335-337),
    the run FAILS when no checkpoint is found under ``cfg.model_dir``; seeded random weights (tests, benchmarks on
    a box without network) have to be asked for explicitly with ``+allow_random_weights=true``;
  * items may be RAW tomograms (``VITDataset(fused=True)``): pre-processing then runs inside the GPU extractor;
  * with several ranks (torchrun) the tomograms of a sample are dealt round-robin to the ranks; every rank writes
    only its own files, no collective is involved;
  * file reads and writes overlap the GPU work (one reader, one writer thread, ``_process_sample``), and
    ``+skip_existing=true`` resumes an interrupted run (both absent from the reference, whose loop is serial).
"""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch

from .. import extract
from ..vit import DinoVisionTransformerB200, build_model
from . import hdf
from .config import instantiate, samples, tomogram_exts
from .shard import process_group, rank_world, shard_round_robin

dino_model = ("facebookresearch/dinov2", "dinov2_vitg14_reg")  # run/dino_features.py:25-28
CHECKPOINT_NAMES = ("dinov2_vitg14_reg4_pretrain.pth", "checkpoints/dinov2_vitg14_reg4_pretrain.pth")


@torch.inference_mode()
def _dino_features(data: torch.Tensor, model: DinoVisionTransformerB200, batch_size: int) -> np.ndarray:
    """Same name, arguments and result as the reference (:31-64): ``np.float16 (C, D, H'/14, W'/14)``.
    ``data`` is either the reference's pre-processed ``[D, 3, H', W']`` float tensor or a raw ``[D, H, W]`` tomogram."""
    if data.dim() == 3:
        return extract.extract_tomogram(data, model, batch_size)
    return extract._dino_features(data, model, batch_size)


def _save_data(data: dict[str, np.ndarray], features: np.ndarray, tomo_name: str, dst_dir: Path, feature_chunk_depth: int = 0) -> None:
    """:109-153: ``data`` stays ``data`` (gzip), every other source dataset goes under ``labels/`` (gzip), a stale
    ``dino_features`` is dropped, the new features are stored uncompressed -- contiguous as the reference does, or with
    ``feature_chunk_depth`` = n > 0 (entry point: ``+feature_chunk_depth=n``) as (C, n, h, w) slabs so
    that a reader of a depth crop touches only its slabs."""
    out: dict[str, np.ndarray] = {}
    for key, arr in data.items():
        if key == "dino_features":
            continue
        out["data" if key == "data" else f"labels/{key}"] = arr
    out["dino_features"] = features
    chunks = None
    if feature_chunk_depth > 0 and features.ndim == 4:
        chunks = {"dino_features": (features.shape[0], int(feature_chunk_depth), features.shape[2], features.shape[3])}
    hdf.write_tomogram(Path(dst_dir) / tomo_name, out, chunks=chunks)


def _read_source(path: Path) -> dict[str, np.ndarray]:
    """:193-200: every dataset of the source file, group members flattened to their own name."""
    return {k.split("/")[-1]: v for k, v in hdf.read_tomogram(path).items()}


def _process_sample(src_dir: Path, dst_dir: Path, csv_dir: Path, model, sample: str, datamodule, batch_size: int,
                    image_dir: Path | None = None, use_sam: bool = False, skip_existing: bool = False, writers: int = 3,
                    feature_chunk_depth: int = 0) -> list[str]:
    """:156-205. Returns the records this rank processed.

    The reference reads, extracts and writes one tomogram after the other (:186-204). At a third of a second of GPU time
    per tomogram the gzip of ``data`` and the 403 MB feature write would dominate, so the three stages overlap here
    (SURVEY.md 8f row f2): a reader thread loads tomogram i+1 and its pass-through datasets while the GPU extracts
    tomogram i, ``writers`` writer threads store the tomograms before it (zlib releases the GIL: the gzip of ``data``, about
    a second per 33 MB of incompressible voxels, is what a single writer cannot keep up with). At most one read and
    ``writers`` writes are in flight (bounded host memory: 0.44 GB each); an error in any thread surfaces at a later
    hand-over or at the end, i.e. inside the same call, where the entry point logs it. ``skip_existing`` (not in the reference) leaves alone result files that already
    hold ``dino_features``, so an interrupted run can be resumed."""
    from concurrent.futures import ThreadPoolExecutor

    tomo_dir, result_dir, csv_file = Path(src_dir) / sample, Path(dst_dir) / sample, Path(csv_dir) / f"{sample}.csv"
    if csv_file.exists():
        import pandas as pd

        records = pd.read_csv(csv_file)["tomo_name"].to_list()
    else:
        records = sorted(f.name for f in tomo_dir.glob("*") if f.suffix in tomogram_exts)
    records = shard_round_robin(records)
    if skip_existing:
        def done(name: str) -> bool:
            try:
                return "dino_features" in hdf.list_keys(result_dir / name)
            except Exception:  # noqa: BLE001 - unreadable / truncated (zipfile.BadZipFile, h5py OSError ...): redo it
                return False

        kept = [r for r in records if not done(r)]
        if len(kept) < len(records):
            logging.info("%s: %d of %d tomograms already have features, skipped", sample, len(records) - len(kept), len(records))
        records = kept
    dataset = instantiate(datamodule["dataset"], data_root=tomo_dir, use_sam=use_sam)(records=records)
    if image_dir is not None:
        logging.warning("export_features=True (PCA colour maps) is outside the hot path and is skipped")

    def load(i: int):
        # one read (and one gunzip) of the source file per tomogram: the extractor input is built from the same
        # ``data`` array that is passed through to the result file
        source = _read_source(tomo_dir / records[i])
        from_array = getattr(dataset, "item_from_array", None)
        item = from_array(source["data"]) if from_array is not None and "data" in source else dataset[i]
        return item, source

    n = len(dataset)
    with ThreadPoolExecutor(max_workers=1, thread_name_prefix="cryovit-read") as reader, \
            ThreadPoolExecutor(max_workers=max(1, writers), thread_name_prefix="cryovit-write") as writer:
        nxt = reader.submit(load, 0) if n else None
        pending: list = []  # writes in flight, oldest first; at most ``writers`` (bounded host memory)
        for i in range(n):
            item, source = nxt.result()
            nxt = reader.submit(load, i + 1) if i + 1 < n else None
            features = _dino_features(item, model, batch_size)
            while len(pending) >= max(1, writers):
                pending.pop(0).result()
            pending.append(writer.submit(_save_data, source, features, records[i], result_dir, feature_chunk_depth))
        for f in pending:
            f.result()
    return records


def checkpoint_names(name: str) -> tuple[str, ...]:
    """File names torch.hub gives the *_reg checkpoints: dinov2_vitg14_reg -> dinov2_vitg14_reg4_pretrain.pth."""
    stem = name.replace("_reg", "_reg4") if name.endswith("_reg") else name
    return (f"{stem}_pretrain.pth", f"checkpoints/{stem}_pretrain.pth", f"{name}.pth")


def load_model(model_dir: Path | str | None, name: str = dino_model[1], allow_random_weights: bool = False) -> DinoVisionTransformerB200:
    """The object the reference gets from torch.hub (:336): the checkpoint under ``model_dir``.

    The reference raises when ``torch.hub.load`` cannot deliver the model (:335-337); so does this: features from
    random weights are 403 MB of noise per tomogram, never a silent default. ``allow_random_weights`` (entry point:
    ``+allow_random_weights=true``) is the explicit override for tests and synthetic benchmarks."""
    tried = []
    if model_dir is not None:
        for rel in dict.fromkeys(checkpoint_names(name) + CHECKPOINT_NAMES):
            p = Path(model_dir) / rel
            tried.append(str(p))
            if p.exists():
                logging.info("loading DINOv2 checkpoint %s", p)
                return build_model(name, state_dict=torch.load(p, map_location="cpu")).cuda().eval()
    if not allow_random_weights:
        raise FileNotFoundError(
            f"no DINOv2 checkpoint for {name}: looked for {tried or 'nothing (model_dir is not set)'}. Download it where "
            "there is network access (torch.hub: facebookresearch/dinov2) and put it under model_dir; "
            "+allow_random_weights=true runs with seeded RANDOM weights instead (tests / synthetic benchmarks only)")
    logging.warning("no DINOv2 checkpoint under %s: +allow_random_weights=true, using seeded RANDOM weights", model_dir)
    return build_model(name).cuda().eval()


def run_trainer(cfg) -> None:
    """:304-350."""
    paths = cfg["paths"]
    data_dir, exp_dir = Path(paths["data_dir"]), Path(paths["exp_dir"])
    src_dir, dst_dir = data_dir / paths["feature_name"], data_dir / paths["tomo_name"]
    csv_dir, image_dir = data_dir / paths["csv_name"], exp_dir / "dino_images"
    sample = cfg.get("sample")
    if sample is not None and not isinstance(sample, str):
        sample = getattr(sample, "name", str(sample))
    sample_names = [sample] if sample is not None else [s for s in samples if (src_dir / s).exists()]
    if cfg.get("use_sam"):
        raise NotImplementedError("use_sam=True (SAM2 image encodings) is outside the B200 hot path")
    rank, world = rank_world()
    with process_group(need_collectives=False):  # binds LOCAL_RANK's GPU; tomograms shard by rank, nothing is exchanged
        _run_samples(cfg, paths, src_dir, dst_dir, csv_dir, image_dir, sample_names, rank, world)


def _run_samples(cfg, paths, src_dir, dst_dir, csv_dir, image_dir, sample_names, rank, world) -> None:
    model_dir = cfg.get("model_dir") or paths.get("model_dir")
    model = load_model(model_dir, cfg.get("dino_variant") or dino_model[1], bool(cfg.get("allow_random_weights", False)))
    import time

    for name in sample_names:
        t0 = time.perf_counter()
        done = _process_sample(src_dir, dst_dir, csv_dir, model, name, cfg["datamodule"], int(cfg["batch_size"]),
                               image_dir if cfg.get("export_features") else None, False, bool(cfg.get("skip_existing", False)),
                               int(cfg.get("writers", 3)), int(cfg.get("feature_chunk_depth", 0)))
        logging.info("rank %d/%d: %d tomograms of %s in %.2f s (files in -> files out, %s container)", rank, world, len(done), name,
                     time.perf_counter() - t0, hdf.backend())
