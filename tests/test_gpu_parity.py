"""End-to-end parity on a B200: the CUDA path (through the C ABI) against the fp32 CPU oracle on the same
seeded inputs and random-init weights, and against the committed golden fixtures.

Tolerances are BASELINE.json's: per-token feature relative error <= 1e-2 and cosine similarity >= 0.999
(both evaluated per patch token over the channel axis; we assert on the worst token and report the mean).
"""
import os
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
REL_TOL, COS_TOL = 1e-2, 0.999


def token_errors(got: torch.Tensor, ref: torch.Tensor):
    """got/ref [..., C] -> (max rel err, mean rel err, min cosine) over tokens."""
    got, ref = got.double().reshape(-1, got.shape[-1]), ref.double().reshape(-1, ref.shape[-1])
    rel = (got - ref).norm(dim=-1) / ref.norm(dim=-1)
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=-1)
    return rel.max().item(), rel.mean().item(), cos.min().item()


def _check(got, ref, what):
    rmax, rmean, cmin = token_errors(got, ref)
    print(f"\n[parity] {what}: rel-err max {rmax:.3e} mean {rmean:.3e}; min cosine {cmin:.6f}")
    assert rmax <= REL_TOL, f"{what}: per-token relative error {rmax:.3e} > {REL_TOL}"
    assert cmin >= COS_TOL, f"{what}: cosine {cmin:.6f} < {COS_TOL}"


@pytest.mark.parametrize("name,shape", [("dinov2_vits14_reg", (2, 3, 392, 392)), ("tiny_swiglu", (2, 3, 56, 84))])
def test_forward_features_vs_oracle_and_golden(cuda_lib, name, shape):
    from cryovit_b200.vit import CONFIGS, DinoVisionTransformerB200, ViTConfig, random_state_dict
    from oracle import dinov2 as odino

    cfg = CONFIGS.get(name) or ViTConfig("tiny_swiglu", 384, 3, 6, "swiglu", 1024)
    sd = random_state_dict(cfg, seed=0)
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(1))
    model = DinoVisionTransformerB200(cfg).load_state_dict(sd).cuda().eval()
    out = model.forward_features(x.cuda())
    got = out["x_norm_patchtokens"].float().cpu()
    ref = odino.forward_features(sd, x, cfg.num_heads)
    _check(got, ref["x_norm_patchtokens"], f"{name} patch tokens vs oracle")
    _check(out["x_norm_clstoken"].float().cpu()[:, None], ref["x_norm_clstoken"][:, None], f"{name} cls vs oracle")
    hf = np.load(GOLD / "dinov2_hf.npz")
    tag = "vits" if name == "dinov2_vits14_reg" else "tiny_swiglu"
    sub = got[:, ::7] if tag == "vits" else got
    _check(sub, torch.from_numpy(hf[f"dinov2_{tag}_hf_patchtokens"]), f"{name} patch tokens vs transformers golden")


@pytest.mark.parametrize("dim,heads,ffn,hidden", [(768, 12, "mlp", 3072), (1024, 16, "mlp", 4096), (1536, 24, "swiglu", 4096)])
@pytest.mark.parametrize("operands", ["bf16", "fp16", "mixed", "mixed-attn"])
def test_every_dinov2_width_vs_oracle(cuda_lib, dim, heads, ffn, hidden, operands):
    """The widths of the other hub entries the reference's ``dino_model`` config can name (ViT-B/14: 768 x 12 heads,
    ViT-L/14: 1024 x 16 heads, ViT-g/14: 1536 x 24 heads with SwiGLU), three blocks each, ragged batch of 3 slices on a
    non-square 4 x 6 patch grid, both operand formats."""
    from cryovit_b200.vit import DinoVisionTransformerB200, ViTConfig, random_state_dict
    from oracle import dinov2 as odino

    cfg = ViTConfig(f"w{dim}", dim, 3, heads, ffn, hidden)
    sd = random_state_dict(cfg, seed=2)
    x = torch.rand(3, 3, 56, 84, generator=torch.Generator().manual_seed(3))
    model = DinoVisionTransformerB200(cfg, operands).load_state_dict(sd).cuda()
    got = model.forward_features(x.cuda())["x_norm_patchtokens"].float().cpu()
    ref = odino.forward_features(sd, x, cfg.num_heads)["x_norm_patchtokens"]
    assert got.shape == ref.shape == (3, 24, dim)
    _check(got, ref, f"width {dim} ({ffn}, {operands} operands) vs oracle")


def test_fused_extract_matches_reference_pipeline(cuda_lib):
    """u8 tomogram -> (fused GPU preproc + ViT-S + write-out) vs oracle preproc -> oracle ViT -> reference layout.
    Includes a non-multiple-of-16 plane (edge-pad path) and a ragged last batch."""
    from cryovit_b200.extract import extract_tomogram
    from cryovit_b200.vit import CONFIGS, build_model, random_state_dict
    from oracle import dinov2 as odino
    from oracle import extract as oextract
    from oracle import preproc as opre

    cfg = CONFIGS["dinov2_vits14_reg"]
    sd = random_state_dict(cfg, seed=0)
    tomo = np.random.default_rng(3).integers(0, 256, size=(5, 100, 120), dtype=np.uint8)
    model = build_model(cfg.name, sd).cuda()
    got = extract_tomogram(tomo, model, batch_size=2)
    ref = oextract.dino_features(opre.dino_transform(opre.load_tomogram(tomo)), odino.OracleDino(sd, cfg.num_heads), 2)
    assert got.dtype == np.float16 and got.shape == ref.shape == (384, 5, 7, 8) and got.flags["C_CONTIGUOUS"]
    g = torch.from_numpy(got.astype(np.float32)).permute(1, 2, 3, 0)
    r = torch.from_numpy(ref.astype(np.float32)).permute(1, 2, 3, 0)
    _check(g, r, "fused extract (ViT-S) vs reference pipeline")


def test_dino_features_mirror_matches_oracle(cuda_lib):
    """The reference-facing _dino_features(data, model, batch_size) signature (seam B3)."""
    from cryovit_b200.extract import _dino_features
    from cryovit_b200.vit import CONFIGS, build_model, random_state_dict
    from oracle import dinov2 as odino
    from oracle import extract as oextract

    cfg = CONFIGS["dinov2_vits14_reg"]
    sd = random_state_dict(cfg, seed=0)
    data = torch.rand(3, 3, 56, 70, generator=torch.Generator().manual_seed(5))
    got = _dino_features(data, build_model(cfg.name, sd).cuda(), batch_size=2)
    ref = oextract.dino_features(data, odino.OracleDino(sd, cfg.num_heads), 2)
    assert got.shape == ref.shape == (384, 3, 4, 5) and got.dtype == np.float16
    _check(torch.from_numpy(got.astype(np.float32)).permute(1, 2, 3, 0),
           torch.from_numpy(ref.astype(np.float32)).permute(1, 2, 3, 0), "_dino_features mirror")


@pytest.fixture(scope="module")
def vitg_sd():
    from cryovit_b200.vit import CONFIGS, random_state_dict

    return random_state_dict(CONFIGS["dinov2_vitg14_reg"], seed=0)


@pytest.fixture(scope="module")
def vitg_oracle_slice(vitg_sd):
    from cryovit_b200.vit import CONFIGS
    from oracle import dinov2 as odino

    x = torch.rand(1, 3, 448, 448, generator=torch.Generator().manual_seed(1))
    return x, odino.forward_features(vitg_sd, x, CONFIGS["dinov2_vitg14_reg"].num_heads)["x_norm_patchtokens"]


@pytest.mark.parametrize("operands", ["mixed-attn", "mixed", "fp16", "bf16"])
def test_vitg_one_slice_vs_oracle(cuda_lib, vitg_sd, vitg_oracle_slice, operands):
    """The headline model (ViT-g/14-reg4, 40 blocks, LayerScale 1.0 random init: the worst case for 16-bit error
    accumulation) on one 448x448 slice against the fp32 oracle evaluated on the host cores: the default operand format
    ("mixed-attn": fp16 norm1 / attention output and the qkv / proj weights, everything else bf16), "mixed" (the FFN input
    side fp16 as well), fp16 everywhere it is bounded, and the all-bf16 alternative -- all within the 1e-2 tolerance; the
    default must stay under 6.5e-3, the two wider formats under 5e-3."""
    from cryovit_b200.vit import CONFIGS, DinoVisionTransformerB200

    cfg = CONFIGS["dinov2_vitg14_reg"]
    x, ref = vitg_oracle_slice
    model = DinoVisionTransformerB200(cfg, operands).load_state_dict(vitg_sd).cuda()
    got = model.forward_features(x.cuda())["x_norm_patchtokens"].float().cpu()
    del model
    torch.cuda.empty_cache()
    _check(got, ref, f"ViT-g one slice vs oracle ({operands} operands)")
    if operands != "bf16":
        assert token_errors(got, ref)[0] <= (6.5e-3 if operands == "mixed-attn" else 5e-3)


def test_vitg_full_size_tomogram_properties(cuda_lib, vitg_sd):
    """BASELINE config 2 at its full size (ViT-g/14-reg4, one 128 x 512 x 512 uint8 tomogram, slice batch 128), where the
    oracle cannot follow, through properties that do not depend on size:
      * a slice's features do not depend on what else is in its batch: slice batch 128 == slice batch 48 (ragged last
        batch of 32), bit for bit -- every token's dot products accumulate over K in the same order whatever M is;
      * slices are independent: permuting the tomogram's slices permutes the features, bit for bit;
      * one slice of the full-size run, re-computed by the fp32 oracle on the host cores, is within the tolerance."""
    from cryovit_b200.extract import extract_tomogram
    from cryovit_b200.vit import CONFIGS, DinoVisionTransformerB200
    from oracle import dinov2 as odino
    from oracle import preproc as opre

    cfg = CONFIGS["dinov2_vitg14_reg"]
    model = DinoVisionTransformerB200(cfg).load_state_dict(vitg_sd).cuda()
    tomo = np.random.default_rng(5).integers(0, 256, size=(128, 512, 512), dtype=np.uint8)
    full = extract_tomogram(tomo, model, batch_size=128)
    assert full.shape == (1536, 128, 32, 32) and full.dtype == np.float16 and np.isfinite(full).all()
    ragged = extract_tomogram(tomo, model, batch_size=48)
    assert np.array_equal(full, ragged), "features depend on the slice batch size"
    perm = np.random.default_rng(6).permutation(128)
    shuffled = extract_tomogram(np.ascontiguousarray(tomo[perm]), model, batch_size=128)
    assert np.array_equal(shuffled, full[:, perm]), "slices are not independent of their position in the batch"
    del model
    torch.cuda.empty_cache()
    k = 77
    x = opre.dino_transform(opre.load_tomogram(tomo[k:k + 1]))  # [1, 3, 448, 448] fp32
    ref = odino.forward_features(vitg_sd, x, cfg.num_heads)["x_norm_patchtokens"][0]  # [1024, 1536]
    got = torch.from_numpy(full[:, k].astype(np.float32)).reshape(1536, 1024).t()
    _check(got, ref.half().float(), "full-size run, slice 77 vs oracle")


def _oracle_fp32_on_gpu(sd_gpu, x, heads):
    """The fp32 oracle evaluated on the GPU in strict fp32 (TF32 off everywhere): same restatement, same arithmetic
    type as on the host cores, fast enough to follow whole batches of ViT-g slices. Checked against the host-core
    evaluation in test_vitg_parity_sweep before it is trusted."""
    from oracle import dinov2 as odino

    flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.get_float32_matmul_precision())
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    try:
        outs = [odino.forward_features(sd_gpu, x[i:i + 4].cuda(), heads)["x_norm_patchtokens"].cpu() for i in range(0, x.shape[0], 4)]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = flags[:2]
        torch.set_float32_matmul_precision(flags[2])
    return torch.cat(outs)


def heavy_tailed_state_dict(cfg, seed, scale):
    """Random init pushed towards what a TRAINED DINOv2 looks like, where random init is benign: LayerScale gammas
    log-uniform in [1e-2, 1] instead of 1.0, a handful of residual-stream channels with ``scale``-fold outliers
    (patch-embed bias, damped by the LayerNorm gains that meet them), and a few FFN hidden units / output channels
    scaled ``scale``-fold (the large-magnitude channels of the *_reg models are born in the FFN)."""
    from cryovit_b200.vit import random_state_dict

    sd = random_state_dict(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    C, Fh = cfg.embed_dim, cfg.hidden
    hot = torch.randperm(C, generator=g)[:6]
    sd["patch_embed.proj.bias"][hot] *= scale
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        for ls in ("ls1.gamma", "ls2.gamma"):
            sd[p + ls] = 10.0 ** (-2.0 * torch.rand(C, generator=g))
        units = torch.randperm(Fh, generator=g)[:4]
        outs = torch.randperm(C, generator=g)[:3]
        if cfg.ffn == "swiglu":
            sd[p + "mlp.w12.weight"][units] *= scale         # silu input of a few hidden units
            sd[p + "mlp.w3.weight"][outs] *= scale           # a few output channels of the FFN
        else:
            sd[p + "mlp.fc1.weight"][units] *= scale
            sd[p + "mlp.fc2.weight"][outs] *= scale
        sd[p + "norm1.weight"][hot] *= 0.1                   # trained models damp their outlier channels in the norms
        sd[p + "norm2.weight"][hot] *= 0.1
    return sd


def _tf32_emulation_on_gpu(sd_gpu, x, heads):
    """The reference's OWN arithmetic (run/dino_features.py:24: every GEMM in TF32) emulated inside the fp32 oracle:
    what the reference itself loses against exact fp32 on the same weights (tests/quant_emulation.py)."""
    from quant_emulation import forward_emulated

    flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.get_float32_matmul_precision())
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    try:
        return forward_emulated(sd_gpu, x.cuda(), heads, "T", "T", "T").cpu()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = flags[:2]
        torch.set_float32_matmul_precision(flags[2])


def _percentiles(got, ref):
    got, ref = got.double().reshape(-1, got.shape[-1]), ref.double().reshape(-1, ref.shape[-1])
    rel = (got - ref).norm(dim=-1) / ref.norm(dim=-1)
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=-1)
    return rel.max().item(), rel.mean().item(), rel.quantile(0.999).item(), cos.min().item()


def test_vitg_parity_sweep(cuda_lib, vitg_sd, vitg_oracle_slice):
    """The ViT-g margin over more than one slice and one set of weights, for the DEFAULT operand
    format (the one bench.py reports): 16 slices of a full-size 128 x 512 x 512 run, three weight seeds on one slice,
    and one heavy-tailed set of weights, each against the fp32 oracle. The oracle is evaluated on the GPU in strict
    fp32 (first checked against its host-core evaluation on the fixture slice). Bar: worst token of the whole set
    <= 8e-3 relative error (the 1e-2 tolerance with 20 % to spare) and cosine >= 0.999."""
    from cryovit_b200.extract import extract_tomogram
    from cryovit_b200.vit import CONFIGS, DinoVisionTransformerB200, random_state_dict
    from oracle import preproc as opre

    cfg = CONFIGS["dinov2_vitg14_reg"]
    x1, ref_cpu = vitg_oracle_slice
    sd_gpu = {k: v.cuda() for k, v in vitg_sd.items()}
    ref_gpu = _oracle_fp32_on_gpu(sd_gpu, x1, cfg.num_heads)
    rmax, _, _, cmin = _percentiles(ref_gpu, ref_cpu)
    print(f"\n[parity] fp32 oracle on the GPU vs on the host cores: rel-err max {rmax:.2e}, min cosine {cmin:.8f}")
    assert rmax < 2e-4, "the GPU evaluation of the oracle is not fp32-faithful"

    worst = {}
    # (a) 16 slices of the full-size run, default format
    model = DinoVisionTransformerB200(cfg).load_state_dict(vitg_sd).cuda()
    assert model.operands == "mixed-attn"
    tomo = np.random.default_rng(5).integers(0, 256, size=(128, 512, 512), dtype=np.uint8)
    full = extract_tomogram(tomo, model, batch_size=128)
    ks = list(range(3, 128, 8))
    x = opre.dino_transform(opre.load_tomogram(tomo[ks]))
    ref = _oracle_fp32_on_gpu(sd_gpu, x, cfg.num_heads).half().float()                       # [16, 1024, 1536]
    got = torch.from_numpy(full[:, ks].astype(np.float32)).reshape(1536, len(ks), 1024).permute(1, 2, 0)
    worst["16 slices of the full-size run"] = _percentiles(got, ref)
    # the all-bf16 alternative on the same slices, for the record (not the default, not asserted at 8e-3)
    del model
    model = DinoVisionTransformerB200(cfg, "bf16").load_state_dict(vitg_sd).cuda()
    got_bf = model.forward_features(x[:4].cuda())["x_norm_patchtokens"].float().cpu()
    bf16_stats = _percentiles(got_bf, ref[:4])
    del model, sd_gpu
    torch.cuda.empty_cache()
    # (b) three more weight seeds, (c) heavy-tailed weights: one slice each. Next to every case: the error of the
    # reference's own TF32 arithmetic on the same weights (emulated in the oracle), the yardstick for ill-conditioned ones.
    cases = [(f"weights seed {s}", random_state_dict(cfg, seed=s), True) for s in (11, 12, 13)]
    cases.append(("heavy-tailed weights x10 (outlier channels, LayerScale 1e-2..1)", heavy_tailed_state_dict(cfg, 21, 10.0), True))
    # x50 saturates the scaled SwiGLU units and sharpens the softmax until near-ties flip: the network itself amplifies
    # every rounding ~5x more than the well-conditioned cases do (TF32 -- the reference's own arithmetic -- is off by
    # 3 % on its worst token there), so this case is held against the reference's own loss, not against the absolute bar
    cases.append(("heavy-tailed weights x50 (ill-conditioned)", heavy_tailed_state_dict(cfg, 21, 50.0), False))
    tf32 = {}
    for name, sd, well_conditioned in cases:
        model = DinoVisionTransformerB200(cfg).load_state_dict(sd).cuda()
        got = model.forward_features(x1.cuda())["x_norm_patchtokens"].float().cpu()
        del model
        sdg = {k: v.cuda() for k, v in sd.items()}
        ref = _oracle_fp32_on_gpu(sdg, x1, cfg.num_heads)
        tf32[name] = _percentiles(_tf32_emulation_on_gpu(sdg, x1, cfg.num_heads), ref)
        del sdg
        torch.cuda.empty_cache()
        (worst if well_conditioned else tf32).setdefault(name + ("" if well_conditioned else " [ours]"), _percentiles(got, ref))
    print("[parity] ViT-g/14-reg4, default operand format (mixed-attn), per-token relative error vs the fp32 oracle:")
    for name, (rmax, rmean, r999, cmin) in worst.items():
        t = tf32.get(name)
        print(f"[parity]   {name}: max {rmax:.3e}  mean {rmean:.3e}  99.9th pct {r999:.3e}  min cosine {cmin:.6f}"
              + (f"   (reference's TF32 arithmetic: max {t[0]:.3e} mean {t[1]:.3e})" if t else ""))
    ill = "heavy-tailed weights x50 (ill-conditioned)"
    ours, ref_tf32 = tf32[ill + " [ours]"], tf32[ill]
    print(f"[parity]   {ill}: ours max {ours[0]:.3e} mean {ours[1]:.3e}; reference's TF32 arithmetic max {ref_tf32[0]:.3e} mean {ref_tf32[1]:.3e}")
    print(f"[parity]   (all-bf16 operands, 4 of the 16 slices: max {bf16_stats[0]:.3e} mean {bf16_stats[1]:.3e})")
    assert max(v[0] for v in worst.values()) <= 8e-3
    assert min(v[3] for v in worst.values()) >= COS_TOL
    # everywhere else the default format loses ~4x what TF32 loses (7- and 10-bit operands against 10-bit ones, 4e-3
    # against 1e-3); the ill-conditioned network must amplify both alike -- anything beyond that would be a kernel defect
    ratios = {n: worst[n][1] / tf32[n][1] for n in worst if n in tf32}
    print(f"[parity]   mean error relative to the reference's TF32 arithmetic: {({k: round(v, 2) for k, v in ratios.items()})}, "
          f"ill-conditioned {ours[1] / ref_tf32[1]:.2f}")
    assert ours[1] <= 1.6 * max(ratios.values()) * ref_tf32[1], "the ill-conditioned case amplifies our roundings more than the reference's"
    assert bf16_stats[0] <= REL_TOL


# ------------------------------------------------------------------------------------------------- head
MASK_AGREEMENT = 0.995  # BASELINE.json: segmentation-mask voxel agreement >= 99.5 % at threshold 0.5


def _head_case(in_ch, D, h, w, seed, spread_bias=False):
    from cryovit_b200.head import CryoVITHeadB200
    from oracle import head as ohead

    sd = ohead.random_state_dict(in_ch, seed=seed)
    if spread_bias:
        # random-init logits sit within ~1e-2 of the output bias, where a 0.5 threshold is a coin toss; rescale the
        # last layer so logits span the clip range and the mask is meaningful (SURVEY.md hard part (g))
        sd["output_layer.2.weight"] = sd["output_layer.2.weight"] * 60.0
    feats = (torch.randn(in_ch, D, h, w, generator=torch.Generator().manual_seed(seed + 1)) * 0.5).half()
    head = CryoVITHeadB200(in_ch).load_state_dict(sd).cuda()
    logits, probs = head.segment_volume(feats.cuda())
    ref_logits = ohead.forward_volume(sd, feats.float()[None])[0, 0]
    return logits.cpu(), probs.cpu(), ref_logits


@pytest.mark.parametrize("in_ch,D,h,w", [(1536, 40, 4, 8), (384, 8, 7, 7), (1536, 6, 2, 3), (1536, 128, 4, 4)])  # last: full depth
def test_head_vs_oracle(cuda_lib, in_ch, D, h, w):
    logits, probs, ref = _head_case(in_ch, D, h, w, seed=3, spread_bias=True)
    assert logits.shape == ref.shape == (D, 16 * h, 16 * w)
    err = (logits - ref).abs()
    got_mask, ref_mask = probs >= 0.5, torch.sigmoid(ref) >= 0.5
    agree = (got_mask == ref_mask).float().mean().item()
    print(f"\n[parity] head in={in_ch} {D}x{h}x{w}: |dlogit| max {err.max():.3e} mean {err.mean():.3e}; "
          f"logit std {ref.std():.3f}; mask agreement {agree:.5f}; positive fraction {ref_mask.float().mean():.3f}")
    assert agree >= MASK_AGREEMENT
    assert err.mean() < 0.05 * ref.std().item() + 1e-3


def test_head_full_size_fitted_vs_oracle(cuda_lib):
    """BASELINE config 4 at its full size: one fp16 (1536, 128, 32, 32) feature volume -> (128, 512, 512) probabilities
    against the fp32 CPU oracle evaluated on the WHOLE volume (about 20 s on the box's host cores). A random-init head
    predicts one class everywhere, so the head is first fitted (B200 training path, 150 AdamW steps at full size) to a
    block mask that the first 64 feature channels carry; both sides then run the same fitted weights. Mask voxel
    agreement >= 99.5 % over all 33.5 M voxels, and the mask must be non-trivial and close to the labels."""
    from cryovit_b200.head import CryoVITHeadB200
    from cryovit_b200.train import CryoVITHeadTrainerB200
    from oracle import head as ohead

    C, D, h, w = 1536, 128, 32, 32
    g = torch.Generator().manual_seed(11)
    coarse = torch.rand(16, 8, 8, generator=g) < 0.3
    cells = coarse.repeat_interleave(8, 0).repeat_interleave(4, 1).repeat_interleave(4, 2)  # (128, 32, 32) feature cells
    feats = torch.randn(C, D, h, w, generator=g) * 0.5
    feats[:64] += cells.float()
    feats = feats.half()
    labels = cells.repeat_interleave(16, 1).repeat_interleave(16, 2).float()  # (128, 512, 512)
    trainer = CryoVITHeadTrainerB200(C, lr=3e-4, state_dict=ohead.random_state_dict(C, seed=5))
    fd, ld = feats.cuda(), labels.cuda()
    losses = [float(trainer.train_step(fd, ld)) for _ in range(150)]
    hsd = {k: v.detach().float().cpu().clone() for k, v in trainer.state_dict().items()}
    del trainer, ld
    torch.cuda.empty_cache()
    head = CryoVITHeadB200(C).load_state_dict(hsd).cuda()
    logits, probs = head.segment_volume(fd)
    logits, probs = logits.cpu(), probs.cpu()
    del head, fd
    torch.cuda.empty_cache()
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ohead.forward_volume(hsd, feats.float()[None])[0, 0]
    assert logits.shape == ref.shape == (D, 16 * h, 16 * w)
    ref_mask, got_mask = torch.sigmoid(ref) >= 0.5, probs >= 0.5
    agree = (got_mask == ref_mask).float().mean().item()
    vs_labels = (ref_mask == (labels > 0.5)).float().mean().item()
    err = (logits - ref).abs()
    print(f"\n[parity] head full size 1536x128x32x32 (fitted, Dice {losses[0]:.3f} -> {losses[-1]:.3f}): |dlogit| max {err.max():.3e} "
          f"mean {err.mean():.3e}; logit std {ref.std():.3f}; mask agreement {agree:.5f}; positive fraction "
          f"{ref_mask.float().mean():.3f}; oracle mask vs labels {vs_labels:.4f}")
    assert losses[-1] < 0.2 and vs_labels > 0.9, "the fit did not produce a meaningful mask"
    assert agree >= MASK_AGREEMENT


def test_head_matches_reference_golden(cuda_lib):
    """Golden logits produced by the reference's own CryoVIT class (tests/golden/reference_src.npz)."""
    from cryovit_b200.head import CryoVITHeadB200
    from oracle import head as ohead

    ref = np.load(GOLD / "reference_src.npz")
    sd = ohead.random_state_dict(1536, seed=3)
    head = CryoVITHeadB200(1536).load_state_dict(sd).cuda()
    for tag in ("a", "b"):
        x = torch.from_numpy(ref[f"head_{tag}_in"])  # fp16 [1, 1536, D, h, w]
        got = head.forward_volume(x.float().cuda())
        want = torch.from_numpy(ref[f"head_{tag}_logits"])
        assert got.shape == want.shape
        err = (got.cpu() - want).abs()
        print(f"\n[parity] head golden {tag}: |dlogit| max {err.max():.3e}, logit range [{want.min():.3f}, {want.max():.3f}]")
        # random-init logits have a spread of ~1e-2; bf16 activations give ~1e-3 absolute error
        assert err.max() < 6e-3
        probs = head.forward(x.float().permute(0, 2, 1, 3, 4).cuda())
        np.testing.assert_allclose(probs.cpu().numpy(), ref[f"head_{tag}_probs"], atol=2e-3)


def test_feature_stream_matches_blocking_extract(cuda_lib):
    """TomogramFeatureStream (copies on their own streams, recycled slots) returns exactly what the blocking
    extract_tomogram returns, for tomograms of different shapes and dtypes submitted back to back."""
    from cryovit_b200.extract import TomogramFeatureStream, extract_tomogram
    from cryovit_b200.vit import CONFIGS, build_model, random_state_dict

    cfg = CONFIGS["dinov2_vits14_reg"]
    model = build_model(cfg.name, random_state_dict(cfg, seed=0)).cuda()
    rng = np.random.default_rng(11)
    tomos = [rng.integers(0, 256, size=(5, 64, 96), dtype=np.uint8), rng.random((3, 50, 70), dtype=np.float32),
             rng.integers(0, 256, size=(7, 32, 48), dtype=np.uint8), rng.integers(0, 256, size=(5, 64, 96), dtype=np.uint8)]
    stream = TomogramFeatureStream(model, batch_size=4, depth=2)
    got = []
    tickets = []
    for t in tomos:
        tickets.append(stream.submit(t))
        if len(tickets) > 1:
            got.append(tickets.pop(0).result().copy())  # valid until `depth` further submits: copy out
    got.append(tickets.pop(0).result().copy())
    for t, g in zip(tomos, got):
        want = extract_tomogram(t, model, batch_size=4)
        assert g.dtype == np.float16 and g.shape == want.shape and np.array_equal(g, want)
