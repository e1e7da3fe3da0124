"""CPU suite (-m "not gpu"): the oracle against the committed golden vectors.

tests/golden/reference_src.npz was produced by executing the reference's own source files
(oracle/make_golden.py); tests/golden/dinov2_hf.npz by transformers' Dinov2WithRegistersModel.
/root/reference is NOT read here -- only the fixtures.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import dinov2 as odino
from oracle import extract as oextract
from oracle import head as ohead
from oracle import metrics as ometrics
from oracle import preproc as opre

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def ref():
    return np.load(GOLD / "reference_src.npz")


@pytest.fixture(scope="module")
def hf():
    return np.load(GOLD / "dinov2_hf.npz")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_preproc_matches_reference(ref, tag):
    u8 = ref[f"preproc_{tag}_in"]
    got = opre.dino_transform(opre.load_tomogram(u8)).numpy()
    want = ref[f"preproc_{tag}_out"]
    assert got.shape == want.shape and got.shape[1] == 3
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
    # three identical channels (the fact the B200 path's channel folding relies on)
    assert np.array_equal(got[:, 0], got[:, 1]) and np.array_equal(got[:, 0], got[:, 2])


def test_feature_layout_matches_reference(ref):
    class Fake:
        def forward_features(self, x):
            B, _, H, W = x.shape
            gh, gw = H // 14, W // 14
            s = x[:, 0, 0, 0].reshape(B, 1, 1)
            p = torch.arange(gh * gw, dtype=torch.float32).reshape(1, -1, 1)
            c = torch.arange(5, dtype=torch.float32).reshape(1, 1, -1)
            return {"x_norm_patchtokens": s * 1000.0 + p + c * 0.125}

    D, H, W = 5, 28, 42
    data = torch.zeros(D, 3, H, W)
    data[:, 0, 0, 0] = torch.arange(D, dtype=torch.float32)
    got = oextract.dino_features(data, Fake(), batch_size=2)
    want = ref["layout_features"]
    assert got.dtype == np.float16 and got.shape == want.shape == (5, D, 2, 3)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_head_matches_reference(ref, tag):
    sd = ohead.random_state_dict(1536, seed=3)
    x = torch.from_numpy(ref[f"head_{tag}_in"]).float()
    logits = ohead.forward_volume(sd, x)
    np.testing.assert_allclose(logits.numpy(), ref[f"head_{tag}_logits"], rtol=0, atol=2e-5)
    probs = ohead.forward(sd, x.permute(0, 2, 1, 3, 4))
    np.testing.assert_allclose(probs.numpy(), ref[f"head_{tag}_probs"], rtol=0, atol=1e-5)
    assert probs.shape[-2:] == (16 * x.shape[-2], 16 * x.shape[-1])


def test_head_state_dict_names_and_size():
    sd = ohead.random_state_dict(1536, seed=0)
    assert sum(v.numel() for v in sd.values()) == 8_401_737
    for k in ("layers.0.weight", "layers.2.layers.0.weight", "layers.2.layers.1.weight", "layers.5.layers.5.bias",
              "output_layer.0.weight", "output_layer.2.bias"):
        assert k in sd


def test_metrics_match_reference(ref):
    y_true, y_pred = torch.from_numpy(ref["metric_y_true"]), torch.from_numpy(ref["metric_y_pred"])
    yp, yt = ometrics.masked_select(y_pred, y_true)
    assert yp.shape == (6, 1)  # two ignore labels dropped
    np.testing.assert_allclose(ometrics.dice_loss(yp, yt).numpy(), ref["metric_dice_loss"], atol=1e-7)
    np.testing.assert_allclose(ometrics.dice_metric(yp, yt).numpy(), ref["metric_dice"], atol=1e-7)
    np.testing.assert_allclose(ometrics.f1_metric(yp, yt).numpy(), ref["metric_f1"], atol=1e-7)
    # hand-computed: preds>=.5 -> [1,0,0,1,1,1] vs [1,0,1,0,1,1]: tp=3 fp=1 fn=1
    assert abs(float(ref["metric_dice"]) - 2 * 3 / (4 + 4 + 1e-3)) < 1e-6


@pytest.mark.parametrize("tag,cfg", [("vits", "dinov2_vits14_reg"), ("tiny_swiglu", None)])
def test_dinov2_oracle_matches_transformers(hf, tag, cfg):
    from cryovit_b200.vit import CONFIGS, ViTConfig, random_state_dict

    if cfg is None:
        c, shape = ViTConfig("tiny_swiglu", 384, 3, 6, "swiglu", 1024), (2, 3, 56, 84)
    else:
        c, shape = CONFIGS[cfg], (2, 3, 392, 392)
    sd = random_state_dict(c, seed=0)
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(1))
    torch.set_num_threads(8)
    out = odino.forward_features(sd, x, c.num_heads)
    pt = out["x_norm_patchtokens"]
    sub = pt[:, ::7] if tag == "vits" else pt
    np.testing.assert_allclose(sub.numpy(), hf[f"dinov2_{tag}_hf_patchtokens"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(out["x_norm_clstoken"].numpy(), hf[f"dinov2_{tag}_hf_cls"], rtol=0, atol=2e-4)


def test_swiglu_hidden_rule():
    from cryovit_b200.vit import CONFIGS

    assert CONFIGS["dinov2_vitg14_reg"].hidden == 4096  # (int(4*1536*2/3)+7)//8*8


def test_fast_gelu_formula():
    """The sm_100a kernels evaluate GELU(erf) as x * sigmoid(x q(x^2)), q an even cubic fitted to logit(Phi(x)) / x
    (csrc/ptx.cuh::gelu_erf). The same float32 arithmetic restated in numpy stays within 2.6e-5 absolute of the float64
    definition everywhere and within 2.1e-4 relative wherever |gelu| >= 0.1 (a ninth of the bf16 rounding step 2^-9 that
    every consumer applies). The Abramowitz-Stegun form it replaced (gelu_erf_as, 5e-7) is checked as well."""
    import math

    import numpy as np

    x = np.concatenate([np.linspace(-12, 12, 400001), [-1e4, -100.0, 100.0, 1e4]]).astype(np.float32)
    f = np.float32
    ref = 0.5 * x.astype(np.float64) * (1.0 + np.vectorize(math.erf)(x.astype(np.float64) / math.sqrt(2.0)))
    z = np.abs(x) * f(0.70710678118654752)
    t = f(1.0) / (f(1.0) + f(0.3275911) * z)
    p = ((((f(1.061405429) * t + f(-1.453152027)) * t + f(1.421413741)) * t + f(-0.284496736)) * t + f(0.254829592)) * t
    e = np.exp2(z * z * f(-1.4426950408889634)).astype(np.float32)
    got_as = f(0.5) * x * (f(1.0) + np.copysign(f(1.0) - p * e, x))
    assert np.abs(got_as - ref).max() < 5e-7
    l2e = f(1.4426950408889634)
    x2 = np.minimum(x * x, f(36.0))
    q = (f(7.030335764e-4) * l2e) * x2 + f(-7.401129204e-2) * l2e
    q = q * x2 + f(-1.5950157686) * l2e
    with np.errstate(over="ignore"):
        e = np.exp2((q * x).astype(np.float32)).astype(np.float32)
    got = (x * (f(1.0) / (f(1.0) + e))).astype(np.float32)
    err = np.abs(got - ref)
    assert err.max() < 2.6e-5, err.max()
    big = np.abs(ref) >= 0.1
    assert (err[big] / np.abs(ref[big])).max() < 2.1e-4
