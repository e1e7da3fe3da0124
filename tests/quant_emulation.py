"""CPU emulation of the 16-bit operand roundings of the B200 ViT path inside the fp32 oracle (TEST INFRASTRUCTURE).

Answers, without a GPU, "which operand format moves the ViT-g parity margin": the fp32 oracle forward is re-run with
every tensor that the CUDA path stores as a 16-bit GEMM / attention operand rounded to that format at the same
point (accumulation stays fp32, as on the tensor cores), and compared with the un-rounded oracle.

    python tests/quant_emulation.py [--depth 40] [--seed 0] [--size 448]

Formats per operand group (b = bf16, h = IEEE fp16, t = TF32 rounding of that group only):  ln (LayerNorm output +
qkv / w12 weights), qkv (q, k, v and the softmax probabilities), attn (attention output + proj weights); the FFN
hidden activations, w3 and the patch embedding are bf16 -- except with ln = "T", which rounds EVERY matmul operand of
the network to TF32 (10 mantissa bits): the arithmetic of the reference's own GPU path
(run/dino_features.py:24 ``torch.set_float32_matmul_precision("high")``), i.e. the yardstick for "as accurate as the
reference is against exact fp32".
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from cryovit_b200.vit import CONFIGS, ViTConfig, random_state_dict  # noqa: E402
from oracle import dinov2 as odino  # noqa: E402

DT = {"b": torch.bfloat16, "h": torch.float16, "f": None}


def tf32(t: torch.Tensor) -> torch.Tensor:
    """Round to TF32 (8 exponent bits, 10 mantissa bits), nearest, ties away from zero (cvt.rna.tf32.f32)."""
    i = t.float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def rnd(t: torch.Tensor, fmt: str) -> torch.Tensor:
    if fmt in ("t", "T"):
        return tf32(t)
    return t if DT[fmt] is None else t.to(DT[fmt]).float()


@torch.no_grad()
def forward_emulated(sd: dict, x: torch.Tensor, heads: int, ln_fmt: str, qkv_fmt: str, attn_fmt: str, eps: float = 1e-6,
                     hid_fmt: str | None = None):
    sd = {k: v.float() for k, v in sd.items()}
    B = x.shape[0]
    C = sd["cls_token"].shape[-1]
    # patch embed: bf16 patches x bf16 folded one-channel weight (three identical channels)
    pe = ln_fmt if ln_fmt in ("f", "T") else "b"
    hf = pe if hid_fmt is None else hid_fmt  # FFN hidden activations and w3 / fc2
    t = F.conv2d(rnd(x.float(), pe), rnd(sd["patch_embed.proj.weight"], pe), sd["patch_embed.proj.bias"], stride=14)
    gh, gw = t.shape[-2:]
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], dim=1)
    t = t + odino.interpolate_pos_encoding(sd["pos_embed"], gh, gw)
    t = torch.cat([t[:, :1], sd["register_tokens"].expand(B, -1, -1), t[:, 1:]], dim=1)
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    hd = C // heads
    for i in range(depth):
        p = f"blocks.{i}."
        ln = rnd(F.layer_norm(t, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps), ln_fmt)
        qkv = rnd(F.linear(ln, rnd(sd[p + "attn.qkv.weight"], ln_fmt), sd[p + "attn.qkv.bias"]), qkv_fmt)
        N = qkv.shape[1]
        q, k, v = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
        s = (q @ k.transpose(-2, -1)) * hd ** -0.5
        pr = rnd(torch.exp(s - s.amax(dim=-1, keepdim=True)), qkv_fmt)  # un-normalised probabilities, as stored in TMEM
        o = (pr @ v) / torch.exp(s - s.amax(dim=-1, keepdim=True)).sum(dim=-1, keepdim=True)
        o = rnd(o.transpose(1, 2).reshape(B, N, C), attn_fmt)
        t = t + sd[p + "ls1.gamma"] * F.linear(o, rnd(sd[p + "attn.proj.weight"], attn_fmt), sd[p + "attn.proj.bias"])
        ln = rnd(F.layer_norm(t, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps), ln_fmt)
        if p + "mlp.w12.weight" in sd:
            x12 = F.linear(ln, rnd(sd[p + "mlp.w12.weight"], ln_fmt), sd[p + "mlp.w12.bias"])
            x1, x2 = x12.chunk(2, dim=-1)
            h = rnd(F.silu(x1) * x2, hf)
            f = F.linear(h, rnd(sd[p + "mlp.w3.weight"], hf), sd[p + "mlp.w3.bias"])
        else:
            h = rnd(F.gelu(F.linear(ln, rnd(sd[p + "mlp.fc1.weight"], ln_fmt), sd[p + "mlp.fc1.bias"])), hf)
            f = F.linear(h, rnd(sd[p + "mlp.fc2.weight"], hf), sd[p + "mlp.fc2.bias"])
        t = t + sd[p + "ls2.gamma"] * f
    xn = F.layer_norm(t, (C,), sd["norm.weight"], sd["norm.bias"], eps)
    return xn[:, 1 + sd["register_tokens"].shape[1]:]


def errors(got, ref):
    got, ref = got.double().reshape(-1, got.shape[-1]), ref.double().reshape(-1, ref.shape[-1])
    rel = (got - ref).norm(dim=-1) / ref.norm(dim=-1)
    cos = F.cosine_similarity(got, ref, dim=-1)
    return rel.max().item(), rel.mean().item(), cos.min().item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--variants", default="bbb,hhh,hbh,hbb,bhb,TTT")
    a = ap.parse_args()
    base = CONFIGS["dinov2_vitg14_reg"]
    cfg = ViTConfig("g", base.embed_dim, a.depth, base.num_heads, base.ffn, base.hidden)
    sd = random_state_dict(cfg, seed=a.seed)
    x = torch.rand(1, 3, a.size, a.size, generator=torch.Generator().manual_seed(1))
    t0 = time.time()
    ref = forward_emulated(sd, x, cfg.num_heads, "f", "f", "f")
    print(f"fp32 reference: {time.time() - t0:.1f} s", flush=True)
    for v in a.variants.split(","):
        got = forward_emulated(sd, x, cfg.num_heads, *v)
        rmax, rmean, cmin = errors(got, ref)
        print(f"ln={v[0]} qkv/P={v[1]} attn/proj={v[2]}: rel-err max {rmax:.3e} mean {rmean:.3e} min cos {cmin:.6f}", flush=True)


if __name__ == "__main__":
    main()
