"""``cryovit.run.eval_model`` for the CryoVIT head (reference run/eval_model.py:103-197 with models/base_model.py:
176-241 ``test_step`` and the ``CsvWriter`` / ``TestPredictionWriter`` callbacks): load ``weights.pt`` of an experiment,
run every test tomogram through the sm_100a head, compute the masked loss and metrics with the fused reduction
kernel, write the per-tomogram metric rows and (optionally) the prediction files. Test tomograms are dealt
round-robin to the ranks of a torchrun launch; every rank writes its own rows (distinct files per sample/split are
touched by one rank at a time only when ranks own disjoint samples -- rank 0 merges otherwise, see ``_merge_rows``)."""
from __future__ import annotations

import logging
from pathlib import Path

import torch

from .._lib import CryovitB200Error
from .callbacks import BatchedModelResult, CsvWriter, TestPredictionWriter
from .config import instantiate
from .shard import process_group, rank_world
from .train_model import build_datamodule, setup_exp_dir


@torch.inference_mode()
def test_step(model, batch) -> BatchedModelResult:
    """base_model.py:176-241 for one collated batch (one tomogram): probabilities over the whole volume, loss and
    metrics over the voxels with label > -1."""
    assert batch.aux_data is not None and "data" in batch.aux_data, "Batch aux_data must contain 'data' key for testing."
    # the reference's opt-in mito mask (base_model.py:192-199): granule predictions scored inside mitochondria only
    mito = batch.aux_data.get("labels/mito")
    out = model._masked_predict(batch, use_mito_mask=mito is not None and len(mito) > 0)
    probs, labels = out["preds_full"], out["labels"]
    losses = {k: float(fn(probs, labels)) for k, fn in model.loss_fns.items()}
    losses["total"] = float(sum(losses.values()))
    metrics = {}
    for k, m in model.metric_fns.items():
        m.reset()
        m.update(probs, labels)
        metrics[k] = float(m.compute())
        m.reset()
    samples, names = batch.metadata.identifiers()
    split = [int(s) for s in batch.metadata.split_id] if batch.metadata.split_id is not None else None
    return BatchedModelResult(num_tomos=batch.num_tomos, samples=samples, tomo_names=names, split_id=split,
                              data=batch.aux_data["data"], label=[t.cpu().numpy() for t in batch.labels],
                              preds=[t.float().cpu().numpy() for t in probs], losses=losses, metrics=metrics)


test_step.__test__ = False  # not a pytest function


def run_trainer(cfg) -> list[BatchedModelResult]:
    """eval_model.py:143-197. Returns this rank's results (metrics only are kept; volumes are dropped after writing)."""
    if cfg.model["_target_"] != "cryovit.models.CryoVIT":
        raise CryovitB200Error(f"model {cfg.model['_target_']} is outside the B200 hot path (CryoVIT head only)")
    with process_group():  # every rank on its own GPU; the group carries the gather of the metric rows
        return _run_trainer(cfg)


def _run_trainer(cfg) -> list[BatchedModelResult]:
    torch.manual_seed(cfg.random_seed)
    cfg = setup_exp_dir(cfg, create=False)
    cfg.paths.results_dir.mkdir(parents=True, exist_ok=True)
    assert cfg.paths.exp_dir.exists(), f"Experiment directory {cfg.paths.exp_dir} does not exist. Run training first."
    ckpt = Path(cfg.ckpt_path) if cfg.get("ckpt_path") else cfg.paths.exp_dir / "weights.pt"
    assert ckpt.exists(), f"{cfg.paths.exp_dir} does not contain a checkpoint."
    if ckpt.suffix != ".pt":
        raise ValueError(f"Unsupported checkpoint format: {ckpt.suffix}. Use .pt or .ckpt files.")
    datamodule = build_datamodule(cfg)
    logging.info("Setup dataset.")
    callbacks = [instantiate(cb) for name, cb in cfg.get("callbacks", {}).items() if name != "rich_progress_bar"]
    model = instantiate({k: v for k, v in cfg.model.items()})
    model.load_state_dict(torch.load(ckpt))
    model.cuda()
    logging.info("Setup model.")
    logging.info("Starting testing.")
    rank, world = rank_world()
    results = []
    for batch in datamodule.test_dataloader():
        res = test_step(model, batch)
        for cb in callbacks:
            if isinstance(cb, CsvWriter) and world > 1:
                continue  # rows are merged by rank 0 below: concurrent read-modify-write of one csv is not safe
            cb.on_test_batch_end(res)
        res.data, res.label, res.preds = [], [], []
        results.append(res)
    if world > 1:
        _merge_rows(results, [cb for cb in callbacks if isinstance(cb, CsvWriter)], rank, world)
    return results


def _merge_rows(results, writers, rank, world) -> None:
    """Several ranks evaluated disjoint tomograms: gather the (small) metric rows on rank 0, which writes the csv."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        raise CryovitB200Error("multi-rank evaluation needs an initialised process group to merge the csv rows")
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    if rank == 0:
        for part in gathered:
            for res in part:
                for w in writers:
                    w.on_test_batch_end(res)
