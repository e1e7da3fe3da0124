"""Generate the golden vectors under tests/golden/ (TEST INFRASTRUCTURE; run in the BUILD container only).

    python oracle/make_golden.py            # needs /root/reference and `transformers`

Two sources pin the oracle:

1. The reference's OWN source files, imported from /root/reference/src with shims for the third-party
   packages that are not installed here (pytorch_lightning, torchmetrics, tensordict, h5py, hydra, omegaconf,
   mrcfile, tifffile, sam2 ...). The shims replace base classes / decorators only; every line of arithmetic
   that runs is the reference's: VITDataset._load_tomogram/_dino_transform, _dino_features (layout + fp16
   cast), CryoVIT/SynthesisBlock, DiceLoss, DiceMetric, F1Metric, TomoDataset._random_crop, collate_fn.
2. transformers' Dinov2WithRegistersModel, an independent implementation of the un-vendored upstream DINOv2
   forward, run with the same (re-keyed) random weights as the oracle.

Inputs are regenerated from seeds by the tests; only outputs (and small inputs) are stored, as .npz.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF_SRC = Path("/root/reference/src")
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))


# ------------------------------------------------------------------------------------------------- shims
def install_shims() -> None:
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    class Callback:
        pass

    class BasePredictionWriter(Callback):
        def __init__(self, *a, **k):
            pass

    class LightningDataModule:
        pass

    pl = mod("pytorch_lightning", LightningModule=LightningModule, Callback=Callback, Trainer=object,
             LightningDataModule=LightningDataModule, seed_everything=lambda *a, **k: None)
    mod("pytorch_lightning.utilities", grad_norm=lambda *a, **k: {})
    mod("pytorch_lightning.callbacks", Callback=Callback, BasePredictionWriter=BasePredictionWriter)
    mod("pytorch_lightning.callbacks.prediction_writer", BasePredictionWriter=BasePredictionWriter)
    mod("pytorch_lightning.utilities.types", STEP_OUTPUT=object)
    pl.utilities = sys.modules["pytorch_lightning.utilities"]

    class Metric(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default.clone()
            setattr(self, name, default.clone())

        def reset(self):
            for k, v in self._defaults.items():
                setattr(self, k, v.clone())

        def forward(self, *a, **k):
            self.update(*a, **k)
            return self.compute()

    mod("torchmetrics", Metric=Metric)

    def tensorclass(cls):
        import dataclasses

        return dataclasses.dataclass(cls)

    mod("tensordict", tensorclass=tensorclass)
    mod("h5py", File=object, Group=object)
    mod("mrcfile")
    mod("tifffile")

    class _CS:
        @staticmethod
        def instance():
            return _CS()

        def store(self, *a, **k):
            pass

    mod("hydra", main=lambda **k: (lambda f: f), compose=None, initialize=None)
    mod("hydra.core")
    mod("hydra.core.config_store", ConfigStore=_CS)
    mod("hydra.utils", instantiate=None)
    mod("omegaconf", MISSING="???", OmegaConf=object, DictConfig=dict)
    sys.path.insert(0, str(REF_SRC))
    # the reference's package __init__ files import everything (sam2, visualisation, ...): register bare packages
    # so that only the modules on the hot path are executed.
    for pkg in ("cryovit", "cryovit.models", "cryovit.datasets", "cryovit.datamodules", "cryovit.run"):
        m = types.ModuleType(pkg)
        m.__path__ = [str(REF_SRC / pkg.replace(".", "/"))]
        sys.modules[pkg] = m


def ref_import(name: str):
    return importlib.import_module(name)


# ------------------------------------------------------------------------------------------------- fixtures
def golden_preproc(out: dict) -> None:
    vd = ref_import("cryovit.datasets.vit_dataset")
    ds = vd.VITDataset(Path("."), False, [])
    rng = np.random.default_rng(7)
    for tag, shape in (("a", (3, 48, 80)), ("b", (2, 50, 70))):  # b exercises the edge-pad path
        u8 = rng.integers(0, 256, size=shape, dtype=np.uint8)
        f = u8.astype(np.float32) / 255.0  # _load_tomogram:86-87
        y = ds._dino_transform(f)
        out[f"preproc_{tag}_in"] = u8
        out[f"preproc_{tag}_out"] = y.numpy()


def golden_layout(out: dict) -> None:
    """_dino_features with a stand-in model whose patch tokens encode (slice, row, col, channel)."""
    torch.Tensor.cuda = lambda self, *a, **k: self  # CPU container: keep the reference's .cuda() call harmless
    sys.modules["cryovit.models"].create_sam_model_from_weights = None
    types_mod = ref_import("cryovit.types")
    cfg_mod = types.ModuleType("cryovit.config")
    cfg_mod.BaseDataModule = object
    cfg_mod.DinoFeaturesConfig = object
    cfg_mod.samples = [s.name for s in types_mod.Sample]
    cfg_mod.tomogram_exts = [".hdf", ".mrc"]
    cfg_mod.DINO_PATCH_SIZE = 14
    sys.modules["cryovit.config"] = cfg_mod
    sys.modules["cryovit.visualization"] = types.ModuleType("cryovit.visualization")
    pca = types.ModuleType("cryovit.visualization.dino_pca")
    pca.export_pca = None
    sys.modules["cryovit.visualization.dino_pca"] = pca
    df = ref_import("cryovit.run.dino_features")

    class Fake:
        def forward_features(self, x):
            B, _, H, W = x.shape
            gh, gw = H // 14, W // 14
            s = x[:, 0, 0, 0].reshape(B, 1, 1)  # slice id carried in the data
            p = torch.arange(gh * gw, dtype=torch.float32).reshape(1, -1, 1)
            c = torch.arange(5, dtype=torch.float32).reshape(1, 1, -1)
            return {"x_norm_patchtokens": s * 1000.0 + p + c * 0.125}

    D, H, W = 5, 28, 42
    data = torch.zeros(D, 3, H, W)
    data[:, 0, 0, 0] = torch.arange(D, dtype=torch.float32)
    feats = df._dino_features(data, Fake(), batch_size=2)  # ragged last batch
    out["layout_features"] = feats


def golden_head_and_metrics(out: dict) -> None:
    types_mod = ref_import("cryovit.types")
    cv = ref_import("cryovit.models.cryovit")
    losses = ref_import("cryovit.models.losses")
    metrics = ref_import("cryovit.models.metrics")
    from oracle import head as ohead

    sd = ohead.random_state_dict(1536, seed=3)
    model = cv.CryoVIT(input_key="dino_features", lr=1e-4, weight_decay=1e-3, losses={}, metrics={})
    model.load_state_dict(sd, strict=True)  # same key names as the reference
    model.eval()
    g = torch.Generator().manual_seed(11)
    for tag, (D, h, w) in (("a", (4, 2, 3)), ("b", (40, 2, 2))):  # b: D > largest dilation (32)
        x = torch.randn(1, 1536, D, h, w, generator=g) * 0.5
        out[f"head_{tag}_in"] = x.half().numpy()  # stored as fp16 (the on-disk dtype); tests up-cast
        with torch.no_grad():
            logits = model.forward_volume(x.half().float())
            probs = model.forward(types.SimpleNamespace(tomo_batch=x.half().float().permute(0, 2, 1, 3, 4)))
        out[f"head_{tag}_logits"] = logits.numpy()
        out[f"head_{tag}_probs"] = probs.numpy()

    # loss / metrics on a hand-sized case, including ignore labels (-1)
    y_true = torch.tensor([[[1.0, 0.0], [1.0, -1.0]], [[0.0, 1.0], [-1.0, 1.0]]]).reshape(1, 2, 2, 2)
    y_pred = torch.tensor([[[0.9, 0.2], [0.4, 0.8]], [[0.6, 0.7], [0.1, 0.5]]]).reshape(1, 2, 2, 2)
    mask = y_true > -1.0
    yp, yt = torch.masked_select(y_pred, mask).view(-1, 1), torch.masked_select(y_true, mask).view(-1, 1)
    out["metric_y_true"], out["metric_y_pred"] = y_true.numpy(), y_pred.numpy()
    out["metric_dice_loss"] = losses.DiceLoss()(yp, yt).numpy()
    dm = metrics.DiceMetric(threshold=0.5)
    out["metric_dice"] = dm(yp, yt).numpy()
    fm = metrics.F1Metric()
    out["metric_f1"] = fm(yp, yt).numpy()


def golden_crop_collate(out: dict) -> None:
    td = ref_import("cryovit.datasets.tomo_dataset")
    import pandas as pd

    ds = td.TomoDataset(pd.DataFrame(), "dino_features", "mito", "split_id", Path("."), train=True)
    rng = np.random.default_rng(5)
    feat = rng.standard_normal((6, 140, 34, 33)).astype(np.float16)
    label = rng.integers(-1, 2, size=(140, 34 * 16, 33 * 16)).astype(np.int8)
    rec = {"input": feat, "label": label}
    np.random.seed(123)
    ds._random_crop(rec)
    out["crop_input_shape"] = np.array(rec["input"].shape)
    out["crop_label_shape"] = np.array(rec["label"].shape)
    out["crop_input_sum"] = np.array(rec["input"].astype(np.float64).sum())
    out["crop_label_sum"] = np.array(rec["label"].astype(np.int64).sum())

    # collate_fn (datamodules/utils.py:13-121): equal depths (the ragged branch pads `data` into `label`, :83-85,
    # and cannot run), batch of 2 -> tomo_batch (B, D, C, h, w) fp32, labels (B, D, H, W) fp32
    types_mod = ref_import("cryovit.types")
    utils = ref_import("cryovit.datamodules.utils")
    items = []
    for i in range(2):
        d = torch.from_numpy(rng.standard_normal((6, 4, 3, 2)).astype(np.float16)).float()
        l = torch.from_numpy(rng.integers(-1, 2, size=(4, 48, 32)).astype(np.int8)).float()
        items.append(types_mod.TomogramData(sample="S", tomo_name=f"t{i}", split_id=i, data=d, label=l, aux_data={}))
    batch = utils.collate_fn(items)
    out["collate_in_data"] = np.stack([it.data.numpy() for it in items])
    out["collate_in_label"] = np.stack([it.label.numpy() for it in items])
    out["collate_tomo_batch"] = batch.tomo_batch.numpy()
    out["collate_labels"] = batch.labels.numpy()
    out["collate_tomo_sizes"] = batch.tomo_sizes.numpy()


HOST_SPLITS = [  # (sample, tomo_name, split_id): the synthetic splits.csv of the host-logic fixtures
    ("A", "a0.hdf", 0), ("A", "a1.hdf", 1), ("A", "a2.hdf", 2), ("A", "a3.hdf", 0), ("B", "b0.hdf", 0), ("B", "b1.hdf", 1),
    ("B", "b2.hdf", 2), ("C", "c0.hdf", 1), ("C", "c1.hdf", 2),
]
HOST_RESULTS = [  # (sample, tomo_name, split_id, dice_metric, f1_metric): rows fed to CsvWriter, incl. a replacement
    ("A", "a0.hdf", 0, 0.5, 0.25), ("A", "a3.hdf", 0, 0.75, 0.5), ("A", "a0.hdf", 0, 0.625, 0.375), ("B", "b0.hdf", None, 0.1, 0.2),
]


def golden_host(tmp: Path) -> dict:
    """Selection rules of the reference's Single/MultiSampleDataModule on a small splits.csv, and the csv text its
    CsvWriter leaves behind (models/callbacks.py:112-206) -- the host logic either side of the head."""
    import pandas as pd

    single = ref_import("cryovit.datamodules.single_sample_datamodule").SingleSampleDataModule
    multi = ref_import("cryovit.datamodules.multi_sample_datamodule").MultiSampleDataModule
    cb = ref_import("cryovit.models.callbacks")
    types_mod = ref_import("cryovit.types")
    split_file = tmp / "splits.csv"
    pd.DataFrame(HOST_SPLITS, columns=["sample", "tomo_name", "split_id"]).to_csv(split_file, index=False)
    out: dict = {"selection": {}}
    cases = {
        "single_A_split0": (single, dict(sample=["A"], split_id=0, split_key="split_id", test_sample=None)),
        "single_A_nosplit_testB": (single, dict(sample=["A"], split_id=None, split_key="split_id", test_sample=["B"])),
        "multi_AB_split1": (multi, dict(sample=["A", "B"], split_id=1, split_key="split_id", test_sample=None)),
        "multi_AB_nosplit_testC": (multi, dict(sample=["A", "B"], split_id=None, split_key="split_id", test_sample=["C"])),
        "multi_AC_split2_testB": (multi, dict(sample=["A", "C"], split_id=2, split_key="split_id", test_sample=["B"])),
    }
    for tag, (cls, kw) in cases.items():
        dm = cls(split_file=split_file, dataset_fn=None, dataloader_fn=None, **kw)
        out["selection"][tag] = {k: [list(map(str, r)) for r in getattr(dm, k + "_df")()[["sample", "tomo_name"]].values.tolist()]
                                 for k in ("train", "val", "test", "predict")}
    w = cb.CsvWriter(tmp / "results")
    for s_, t_, sid, dice, f1 in HOST_RESULTS:
        res = types_mod.BatchedModelResult(num_tomos=1, samples=[s_], tomo_names=[t_], split_id=None if sid is None else [sid],
                                           data=[], label=[], preds=[], losses={}, metrics={"dice_metric": dice, "f1_metric": f1},
                                           aux_data=None)
        w.on_test_batch_end(None, None, res, None, 0)
    out["csv"] = {f.name: f.read_text() for f in sorted((tmp / "results").glob("*.csv"))}
    return out


CONFIG_CASES = {
    "dino_features": ("dino_features", ["sample=Q18", "batch_size=64"]),
    "train_multi": ("train_model", ["model=cryovit", "+experiments=multi_mito", "datamodule.sample=[Q18,Q53]",
                                    "datamodule.split_id=3", "label_key=mito"]),
    "train_single": ("train_model", ["model=cryovit", "datamodule=single", "datamodule.sample=Q18", "label_key=mito"]),
    "eval_multi": ("eval_model", ["model=cryovit", "datamodule=multi", "datamodule.sample=[Q18]",
                                  "datamodule.test_sample=[Q53]", "label_key=mito"]),
}
CONFIG_PATHS = ["paths.model_dir=/m", "paths.data_dir=/d", "paths.exp_dir=/e"]


def golden_configs() -> dict:
    """The reference's OWN YAML tree (src/cryovit/configs/**), composed for the command lines of the three entry points:
    what the built-in tree of cryovit_b200/host/config_tree.py has to reproduce key for key. Composed with this repo's
    Hydra-subset composer pointed at the reference's directory (Hydra is not installed here); the schema defaults the
    reference supplies from ConfigStore dataclasses (config.py:30-156) are therefore absent on this side and show up
    as ``???`` or missing keys, which the test skips."""
    from cryovit_b200.host.config import compose

    ref_dir = REF_SRC / "cryovit" / "configs"
    return {k: {"config_name": name, "overrides": ov + CONFIG_PATHS, "composed": compose(name, ov + CONFIG_PATHS, config_dir=ref_dir)}
            for k, (name, ov) in CONFIG_CASES.items()}


def golden_dinov2(out: dict) -> None:
    """Oracle restatement vs transformers' independent Dinov2WithRegisters on identical weights."""
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    from cryovit_b200.vit import CONFIGS, ViTConfig, random_state_dict
    from oracle import dinov2 as od

    def to_hf(sd, cfg):
        C = cfg.embed_dim
        hf = {
            "embeddings.cls_token": sd["cls_token"], "embeddings.mask_token": sd["mask_token"],
            "embeddings.register_tokens": sd["register_tokens"], "embeddings.position_embeddings": sd["pos_embed"],
            "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
            "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
            "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"],
        }
        for i in range(cfg.depth):
            p, q = f"blocks.{i}.", f"encoder.layer.{i}."
            wq, wk, wv = sd[p + "attn.qkv.weight"].split(C, dim=0)
            bq, bk, bv = sd[p + "attn.qkv.bias"].split(C, dim=0)
            hf.update({
                q + "norm1.weight": sd[p + "norm1.weight"], q + "norm1.bias": sd[p + "norm1.bias"],
                q + "norm2.weight": sd[p + "norm2.weight"], q + "norm2.bias": sd[p + "norm2.bias"],
                q + "attention.attention.query.weight": wq, q + "attention.attention.query.bias": bq,
                q + "attention.attention.key.weight": wk, q + "attention.attention.key.bias": bk,
                q + "attention.attention.value.weight": wv, q + "attention.attention.value.bias": bv,
                q + "attention.output.dense.weight": sd[p + "attn.proj.weight"],
                q + "attention.output.dense.bias": sd[p + "attn.proj.bias"],
                q + "layer_scale1.lambda1": sd[p + "ls1.gamma"], q + "layer_scale2.lambda1": sd[p + "ls2.gamma"],
            })
            if cfg.ffn == "swiglu":
                hf.update({q + "mlp.weights_in.weight": sd[p + "mlp.w12.weight"], q + "mlp.weights_in.bias": sd[p + "mlp.w12.bias"],
                           q + "mlp.weights_out.weight": sd[p + "mlp.w3.weight"], q + "mlp.weights_out.bias": sd[p + "mlp.w3.bias"]})
            else:
                hf.update({q + "mlp.fc1.weight": sd[p + "mlp.fc1.weight"], q + "mlp.fc1.bias": sd[p + "mlp.fc1.bias"],
                           q + "mlp.fc2.weight": sd[p + "mlp.fc2.weight"], q + "mlp.fc2.bias": sd[p + "mlp.fc2.bias"]})
        return hf

    cases = {
        "vits": (CONFIGS["dinov2_vits14_reg"], (2, 3, 392, 392)),
        "tiny_swiglu": (ViTConfig("tiny_swiglu", 384, 3, 6, "swiglu", 1024), (2, 3, 56, 84)),
    }
    for tag, (cfg, shape) in cases.items():
        sd = random_state_dict(cfg, seed=0)
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(1))
        ours = od.forward_features(sd, x, cfg.num_heads)
        hcfg = Dinov2WithRegistersConfig(
            hidden_size=cfg.embed_dim, num_hidden_layers=cfg.depth, num_attention_heads=cfg.num_heads,
            mlp_ratio=4, use_swiglu_ffn=cfg.ffn == "swiglu", num_register_tokens=cfg.num_register_tokens,
            image_size=518, patch_size=14, layerscale_value=1.0, layer_norm_eps=1e-6, qkv_bias=True,
            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, drop_path_rate=0.0)
        hf = Dinov2WithRegistersModel(hcfg).eval()
        if cfg.ffn == "swiglu":  # HF derives the hidden size from mlp_ratio; check it matches ours
            assert hf.encoder.layer[0].mlp.weights_in.out_features == 2 * cfg.hidden, (hf.encoder.layer[0].mlp.weights_in.out_features, cfg.hidden)
        missing, unexpected = hf.load_state_dict(to_hf(sd, cfg), strict=False)
        assert not unexpected and all("pooler" in m for m in missing), (missing, unexpected)
        with torch.no_grad():
            hs = hf(pixel_values=x).last_hidden_state  # final-LayerNorm'ed tokens [B, 1+R+Np, C]
        R = cfg.num_register_tokens
        theirs = hs[:, 1 + R:]
        diff = (ours["x_norm_patchtokens"] - theirs).abs().max().item()
        print(f"dinov2 {tag}: oracle vs transformers max abs diff {diff:.3e} (ref max {theirs.abs().max().item():.3f})")
        assert diff < 2e-4, diff
        sub = theirs[:, ::7] if tag == "vits" else theirs
        out[f"dinov2_{tag}_hf_patchtokens"] = sub.numpy().astype(np.float32)
        out[f"dinov2_{tag}_hf_cls"] = hs[:, 0].numpy()


def main() -> None:
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    install_shims()
    ref_out: dict = {}
    golden_preproc(ref_out)
    golden_head_and_metrics(ref_out)
    golden_crop_collate(ref_out)
    golden_layout(ref_out)
    np.savez_compressed(GOLD / "reference_src.npz", **ref_out)
    import json
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        host = golden_host(Path(td))
    host["composed_configs"] = golden_configs()
    (GOLD / "reference_host.json").write_text(json.dumps(host, indent=1))
    hf_out: dict = {}
    golden_dinov2(hf_out)
    np.savez_compressed(GOLD / "dinov2_hf.npz", **hf_out)
    for f in sorted(GOLD.glob("*.npz")):
        print(f.name, f.stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    main()
